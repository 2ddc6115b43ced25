#!/bin/bash
# N GPUs: the device-driven peer exchange against the host-driven NCCL exchange
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
shift
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 10 --warmup 3 "$@"; }
run > gpurun_out/d2_peer_n$N.json 2> gpurun_out/d2_peer_n$N.err
echo "peer rc $?"
run --exchange nccl > gpurun_out/d2_nccl_n$N.json 2> gpurun_out/d2_nccl_n$N.err
echo "nccl rc $?"
tail -5 gpurun_out/d2_peer_n$N.err; head -c 1500 gpurun_out/d2_peer_n$N.json; echo; head -c 600 gpurun_out/d2_nccl_n$N.json
