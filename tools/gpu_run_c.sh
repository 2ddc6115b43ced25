#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shards.py tests/test_gpu_parity.py tests/test_gpu_compat.py -m gpu -q -x --durations=8 --deselect tests/test_gpu_parity.py::test_branch_free_sqrt_and_division_are_ieee > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
timeout 600 python -m pytest tests/test_gpu_ref_oracle.py -m gpu -q -x --durations=8 -k "0.5-64-2 or 1.0-32-2 or table_functor or fullsize or mandelbulb" > gpurun_out/c_pytest_ref.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest_ref.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_c3.json 2> gpurun_out/c_bench_c3.err
SDM_NO_BINS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_c3_nobins.json 2> gpurun_out/c_bench_c3_nobins.err
for c in 32 128 256; do
  SDM_BIN_CHUNK=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_c3_binchunk$c.json 2> gpurun_out/c_bench_c3_binchunk$c.err
done
SDM_SLACK=0.6 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_c3_slack0.6.json 2> gpurun_out/c_bench_c3_slack0.6.err
timeout 300 python bench.py --workload sd_obj_1024 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_sdobj1024.json 2> gpurun_out/c_bench_sdobj1024.err
tail -15 gpurun_out/c_pytest.log; tail -12 gpurun_out/c_pytest_ref.log; head -c 300 gpurun_out/c_bench_c3.json
