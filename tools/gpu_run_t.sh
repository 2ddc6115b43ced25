#!/bin/bash
# one GPU: late weld-key insert (k_vertex_normals) and pipelined gather (k_orient) variants on the default workload
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=$PWD/bevy-signed-distance-mesh-generation_b200
for v in base lw pipe lwpipe; do
  if [ $v = base ]; then L=$P/libsdfmesh.so; else L=$P/libsdfmesh_$v.so; fi
  SDM_LIB=$L timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench_$v.json 2> gpurun_out/t_bench_$v.err
done
python - <<PY
import json
for v in "base lw pipe lwpipe".split():
    try:
        d=json.loads(open(f"gpurun_out/t_bench_{v}.json").read().strip().splitlines()[-1]); k=d['kernel_ms']
        print(v, 'ms=%.3f'%d['ms_per_step'], {a:round(k[a],3) for a in ('k_refine','k_edges','k_project','k_vertex_normals','k_orient')}, d['mesh_fnv'])
    except Exception as e: print(v,'ERR',e)
PY
exit 0
