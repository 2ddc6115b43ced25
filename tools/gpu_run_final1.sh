#!/bin/bash
# one GPU, final build: default bench (+ CPU baseline), reference arm, the other single-GPU configurations, ncu launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/fin_bench_c3.json 2> gpurun_out/fin_bench_c3.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/fin_bench_c3_reference.json 2> gpurun_out/fin_bench_c3_reference.err
timeout 300 python bench.py --workload c2_sd_obj_512 --no-cpu-baseline > gpurun_out/fin_bench_c2.json 2> gpurun_out/fin_bench_c2.err
timeout 300 python bench.py --workload c1_sphere_box_128 --no-cpu-baseline > gpurun_out/fin_bench_c1.json 2> gpurun_out/fin_bench_c1.err
timeout 300 python bench.py --workload c4_mandelbulb_2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/fin_bench_c4.json 2> gpurun_out/fin_bench_c4.err
timeout 400 python bench.py --workload c5_animated_1024 > gpurun_out/fin_bench_c5.json 2> gpurun_out/fin_bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/fin_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fin_ncu_launches.log 2>&1
python - <<PY
import json
for f in ("fin_bench_c3","fin_bench_c3_reference","fin_bench_c2","fin_bench_c1","fin_bench_c4","fin_bench_c5"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, 'ms=%.3f'%d['ms_per_step'], 'value=%.4g'%d['value'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'roofline', (d.get('roofline') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('latency_ms'), d.get('clocks'))
    except Exception as e:
        print(f,'ERR',e); print(open(f"gpurun_out/{f}.err").read()[-800:])
PY
wc -l gpurun_out/fin_launches_c3.csv
exit 0
