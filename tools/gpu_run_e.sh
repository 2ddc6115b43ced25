#!/bin/bash
# N GPUs, peer exchange only (+ optional extra args)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
TAG=${2:-e}
shift; shift
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_peer_n$N.json 2> gpurun_out/${TAG}_peer_n$N.err
echo "rc $?"
tail -3 gpurun_out/${TAG}_peer_n$N.err; python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_peer_n$N.json").read().strip().splitlines()[-1])
print('ms=%.2f gpu=%.2f e2e=%.2f'%(d['ms_per_step'], d['gpu_ms_per_step'], d['e2e']['ms_per_step']), d.get('exchange'), d['mesh_fnv'])
print({a:round(b,3) for a,b in d['kernel_ms'].items() if b>0.02})
PY
