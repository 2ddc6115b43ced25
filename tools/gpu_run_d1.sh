#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=bevy-signed-distance-mesh-generation_b200
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shards.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_branch_free_sqrt_and_division_are_ieee > gpurun_out/d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_c3.json 2> gpurun_out/d_bench_c3.err
for v in split5 split6 split8; do
  SDM_LIB=$PWD/$P/libsdfmesh_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_c3_$v.json 2> gpurun_out/d_bench_c3_$v.err
done
# the split build must give the same bytes: one parity test through it
SDM_LIB=$PWD/$P/libsdfmesh_split6.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "remesh_matches_oracle or fullsize" > gpurun_out/d_pytest_split.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest_split.log
tail -4 gpurun_out/d_pytest.log; tail -4 gpurun_out/d_pytest_split.log
for f in gpurun_out/d_bench_c3*.json; do echo $f; head -c 400 $f | grep -o '"ms_per_step": [0-9.]*'; done
