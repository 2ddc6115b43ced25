#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=12 --deselect tests/test_gpu_parity.py::test_branch_free_sqrt_and_division_are_ieee > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_c3.json 2> gpurun_out/b_bench_c3.err
SDM_NO_LATTICE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_c3_nolattice.json 2> gpurun_out/b_bench_c3_nolattice.err
SDM_SLACK=0.75 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_c3_slack0.75.json 2> gpurun_out/b_bench_c3_slack0.75.err
timeout 300 python tools/perf_probe.py many1024_1024 > gpurun_out/b_probe.log 2>&1 &&
SDM_PROFILE_LAST=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_refine|k_project|k_vertex_normals|k_orient|k_edges|k_emit' -o gpurun_out/b_prof python tools/perf_probe.py many1024_1024 > gpurun_out/b_ncu.log 2>&1
tail -25 gpurun_out/b_pytest.log; cat gpurun_out/b_probe.log | tail -30
