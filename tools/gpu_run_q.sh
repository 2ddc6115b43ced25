#!/bin/bash
# one GPU: six-sample orientation test - whole default GPU suite, default bench with and without it, sd_obj @1024^3, C5
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/q_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_c3.json 2> gpurun_out/q_bench_c3.err
SDM_NO_QUICK_ORIENT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_c3_noquick.json 2> gpurun_out/q_bench_c3_noquick.err
timeout 300 python bench.py --workload sd_obj_1024 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_sdobj.json 2> gpurun_out/q_bench_sdobj.err
tail -4 gpurun_out/q_pytest.log
python - <<PY
import json
for f in ("q_bench_c3","q_bench_c3_noquick","q_bench_sdobj"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); k=d['kernel_ms']
        print(f,'ms=%.3f e2e=%.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), {a:round(v,3) for a,v in k.items() if v > 0.05}, d['mesh_fnv'], 'pending', d.get('orient_pending_triangles'), 'tris', d['triangles'])
    except Exception as e:
        print(f,'ERR',e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
exit 0
