#!/bin/bash
# one GPU: the whole default GPU suite, the default bench, C5 (outlier re-measured with the adaptive tail ball)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/w_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench_c3.json 2> gpurun_out/w_bench_c3.err
timeout 400 python bench.py --workload c5_animated_1024 > gpurun_out/w_bench_c5.json 2> gpurun_out/w_bench_c5.err
tail -4 gpurun_out/w_pytest.log
python - <<PY
import json
d=json.loads(open("gpurun_out/w_bench_c3.json").read().strip().splitlines()[-1]); k=d['kernel_ms']
print('c3 ms=%.3f e2e=%s'%(d['ms_per_step'], d['e2e']), {a:round(v,3) for a,v in k.items()}, d['mesh_fnv']['indices'], d['roofline'])
d=json.loads(open("gpurun_out/w_bench_c5.json").read().strip().splitlines()[-1])
print('c5', d['ms_per_step'], d.get('latency_ms'), d.get('newton'), d.get('outliers'))
PY
