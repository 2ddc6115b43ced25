#!/bin/bash
# N GPUs: default workload (C3) and, optionally, C4 through the peer exchange
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
TAG=${2:-s}
WITH_C4=${3:-0}
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@"; }
run --steps 20 --warmup 3 > gpurun_out/${TAG}_c3_n$N.json 2> gpurun_out/${TAG}_c3_n$N.err
echo "c3 rc $?"
if [ "$WITH_C4" = "1" ]; then
  run --steps 5 --warmup 3 --workload c4_mandelbulb_2048 > gpurun_out/${TAG}_c4_n$N.json 2> gpurun_out/${TAG}_c4_n$N.err
  echo "c4 rc $?"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_c?_n$N.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,'ms=%.3f gpu=%.3f e2e=%.3f'%(d['ms_per_step'], d['gpu_ms_per_step'], d['e2e']['ms_per_step']), d.get('exchange','')[:20], d['mesh_fnv'])
        for r in d.get('per_rank') or []:
            k=r['kernel_ms']; comp=sum(v for a,v in k.items() if not a.startswith('peer'))
            print('  ',r['rank'], r['finest_voxels'], 'gpu',r['gpu_ms'], 'compute %.2f'%comp, {a:v for a,v in k.items() if a.startswith('peer')})
    except Exception as e:
        print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
