#!/bin/bash
# tuning variants of the library (SDM_LIB) on the default workload + the compat render test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=$PWD/bevy-signed-distance-mesh-generation_b200
timeout 300 python -m pytest tests/test_gpu_compat.py -m gpu -q -x > gpurun_out/v_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/v_pytest.log
for v in base pf lu4 lu1 p4 p6 n5 n7 r2 r4; do
  if [ $v = base ]; then L=$P/libsdfmesh.so; else L=$P/libsdfmesh_$v.so; fi
  SDM_LIB=$L timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_$v.json 2> gpurun_out/v_bench_$v.err
done
tail -3 gpurun_out/v_pytest.log
python - <<PY
import json
for v in "base pf lu4 lu1 p4 p6 n5 n7 r2 r4".split():
    try:
        d=json.loads(open(f"gpurun_out/v_bench_{v}.json").read().strip().splitlines()[-1]); k=d['kernel_ms']
        print(v, 'ms=%.3f'%d['ms_per_step'], {a:round(k[a],3) for a in ('k_refine','k_edges','k_project','k_vertex_normals','k_orient')}, d['mesh_fnv']['indices'])
    except Exception as e: print(v,'ERR',e)
PY
