#!/bin/bash
# first GPU pass of round 2: parity tests (all), then the default bench with A/B switches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; free -g | head -2 >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_ref_oracle.py::test_c3_fullsize_matches_reference_kernels > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c3.json 2> gpurun_out/a_bench_c3.err
SDM_NO_LISTS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c3_nolists.json 2> gpurun_out/a_bench_c3_nolists.err
SDM_NO_LATTICE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c3_nolattice.json 2> gpurun_out/a_bench_c3_nolattice.err
for s in 0.5 0.75 1.5; do
  SDM_SLACK=$s timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c3_slack$s.json 2> gpurun_out/a_bench_c3_slack$s.err
done
for c in 64 128 512; do
  SDM_PROJ_CHUNK=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c3_chunk$c.json 2> gpurun_out/a_bench_c3_chunk$c.err
done
timeout 900 python -m pytest tests/test_gpu_ref_oracle.py::test_c3_fullsize_matches_reference_kernels -q -x > gpurun_out/a_pytest_fullsize.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest_fullsize.log
tail -3 gpurun_out/a_pytest.log; tail -3 gpurun_out/a_pytest_fullsize.log; head -c 600 gpurun_out/a_bench_c3.json
