#!/bin/bash
# one GPU: the tail kernel's list rebuild, serial (old) vs candidate-parallel (new), on the two slow frames of configs[4]; parity tests; C4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=$PWD/bevy-signed-distance-mesh-generation_b200
SDM_LIB=$P/libsdfmesh_serialtail.so timeout 300 python tools/c5_frame_probe.py 254 329 > gpurun_out/x_probe_serial.log 2>&1
timeout 300 python tools/c5_frame_probe.py 254 329 100 > gpurun_out/x_probe_new.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_oracle.py -m gpu -q -x > gpurun_out/x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/x_pytest.log
timeout 300 python bench.py --workload c4_mandelbulb_2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench_c4.json 2> gpurun_out/x_bench_c4.err
echo serial; cat gpurun_out/x_probe_serial.log | tail -5; echo new; tail -7 gpurun_out/x_probe_new.log; tail -3 gpurun_out/x_pytest.log
python - <<PY
import json
d=json.loads(open("gpurun_out/x_bench_c4.json").read().strip().splitlines()[-1]); k=d['kernel_ms']
print('c4 ms=%.3f'%d['ms_per_step'], {a:round(v,2) for a,v in k.items() if v>1}, d['mesh_fnv'])
PY
