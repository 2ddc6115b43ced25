#!/bin/bash
# one GPU: slow frames of configs[4] again, the whole default GPU suite, default bench, C5 bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/c5_frame_probe.py 254 329 > gpurun_out/z_probe.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/z_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/z_bench_c3.json 2> gpurun_out/z_bench_c3.err
timeout 400 python bench.py --workload c5_animated_1024 > gpurun_out/z_bench_c5.json 2> gpurun_out/z_bench_c5.err
tail -4 gpurun_out/z_probe.log; tail -4 gpurun_out/z_pytest.log
python - <<PY
import json
d=json.loads(open("gpurun_out/z_bench_c3.json").read().strip().splitlines()[-1]); k=d['kernel_ms']
print('c3 ms=%.3f e2e=%.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), {a:round(v,3) for a,v in k.items()}, d['mesh_fnv']['indices'], d['roofline']['frac'])
d=json.loads(open("gpurun_out/z_bench_c5.json").read().strip().splitlines()[-1])
print('c5', d['ms_per_step'], d.get('latency_ms'), d.get('newton'), d.get('outliers'))
PY
exit 0
