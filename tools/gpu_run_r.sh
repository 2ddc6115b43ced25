#!/bin/bash
# one GPU: corner signs inherited by k_refine - parity tests on the default build, then the default bench for each batch-shape variant
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=$PWD/bevy-signed-distance-mesh-generation_b200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_oracle.py tests/test_gpu_shards.py -m gpu -q -x > gpurun_out/r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r_pytest.log
SDM_NO_CORNER_SIGNS=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_off.json 2> gpurun_out/r_bench_off.err
for v in b10m3 b10m2 b7m3 b5m3 b5m4; do
  SDM_LIB=$P/libsdfmesh_$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_$v.json 2> gpurun_out/r_bench_$v.err
done
tail -4 gpurun_out/r_pytest.log
python - <<PY
import json
for v in "off b10m3 b10m2 b7m3 b5m3 b5m4".split():
    try:
        d=json.loads(open(f"gpurun_out/r_bench_{v}.json").read().strip().splitlines()[-1]); k=d['kernel_ms']
        print(v, 'ms=%.3f'%d['ms_per_step'], {a:round(k[a],3) for a in ('k_refine','k_refine_emit','k_edges','k_project','k_vertex_normals','k_orient')}, d['mesh_fnv']['indices'], d['prim_point_evals_per_step'][0])
    except Exception as e: print(v,'ERR',e)
PY
exit 0
