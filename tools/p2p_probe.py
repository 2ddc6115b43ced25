"""torchrun probe: time of a grouped NCCL send/recv of 3 buffers (100+100+200 MB) from rank 1 to rank 0 under the current NCCL env."""
import os, time, torch, torch.distributed as dist
os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
sizes = [25_000_000 // (world - 1) * 1, 25_000_000 // (world - 1), 50_000_000 // (world - 1)]   # floats per sender
bufs = {r: [torch.empty(n, dtype=torch.float32, device="cuda") for n in sizes] for r in (range(1, world) if rank == 0 else [rank])}
def step():
    ops = []
    if rank == 0:
        for r in range(1, world):
            ops += [dist.P2POp(dist.irecv, b, r) for b in bufs[r]]
    else:
        ops += [dist.P2POp(dist.isend, b, 0) for b in bufs[rank]]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    torch.cuda.current_stream().synchronize()
for _ in range(3): step()
dist.barrier(); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10): step()
dist.barrier(); torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
if rank == 0:
    total = sum(sizes) * 4 * (world - 1)
    print(f"env MINP2P={os.environ.get('NCCL_MIN_P2P_NCHANNELS')} MAXP2P={os.environ.get('NCCL_MAX_P2P_NCHANNELS')} NCHPP={os.environ.get('NCCL_NCHANNELS_PER_NET_PEER')}: {dt*1e3:.3f} ms for {total/1e6:.0f} MB -> {total/dt/1e9:.0f} GB/s")
dist.destroy_process_group()
