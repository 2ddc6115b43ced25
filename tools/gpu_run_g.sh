#!/bin/bash
# one GPU: render test again, C5 with outlier diagnostics
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_render.py tests/test_gpu_shards.py -m gpu -q -x > gpurun_out/g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_pytest.log
timeout 300 python bench.py --workload c5_animated_1024 > gpurun_out/g_bench_c5.json 2> gpurun_out/g_bench_c5.err
tail -5 gpurun_out/g_pytest.log; python - <<PY
import json
d=json.loads(open("gpurun_out/g_bench_c5.json").read().strip().splitlines()[-1])
print(d['latency_ms'], d['newton'], d['outliers'])
PY
