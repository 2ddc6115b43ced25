#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_shards.py tests/test_gpu_parity.py tests/test_gpu_compat.py -m gpu -q -x --durations=6 --deselect tests/test_gpu_parity.py::test_branch_free_sqrt_and_division_are_ieee > gpurun_out/f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_c3.json 2> gpurun_out/f_bench_c3.err
timeout 300 python bench.py --workload c5_animated_1024 > gpurun_out/f_bench_c5.json 2> gpurun_out/f_bench_c5.err
timeout 300 python bench.py --workload c4_mandelbulb_2048 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_c4.json 2> gpurun_out/f_bench_c4.err
tail -14 gpurun_out/f_pytest.log; for f in f_bench_c3 f_bench_c5 f_bench_c4; do head -c 700 gpurun_out/$f.json; echo; tail -2 gpurun_out/$f.err; done
