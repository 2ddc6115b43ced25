"""GPU probe: where does the end-to-end step go?  (D2H bandwidth, per-phase host times of the e2e loop, overlap check)"""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import bsdmg_b200
from bsdmg_b200 import scenes

scene = scenes.SCENES["many1024"]()
h = bsdmg_b200.CudaHandler(0, scene)
bb, init, levels = 5.0, 64, 4
for _ in range(3):
    m = h.remesh(bb, init, levels, download=False)
V, T = int(m.vertex_count), int(m.triangle_count)
print("V", V, "T", T, "gpu_ms", h.stats()["last_gpu_ms"])
bufs = [(torch.empty((V, 3), dtype=torch.float32).pin_memory(), torch.empty((V, 3), dtype=torch.float32).pin_memory(),
         torch.empty((T, 3), dtype=torch.int32).pin_memory()) for _ in range(2)]
nbytes = V * 24 + T * 12
# 1. pure download bandwidth
for rep in range(3):
    t = time.perf_counter()
    h.download_into_async(m, *(b.data_ptr() for b in bufs[0]))
    h.download_wait()
    dt = time.perf_counter() - t
    print(f"download alone: {dt*1e3:.2f} ms  {nbytes/dt/1e9:.1f} GB/s")
# torch's own copy for comparison
d = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
hp = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
torch.cuda.synchronize()
for rep in range(2):
    t = time.perf_counter(); hp.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"torch D2H pinned: {dt*1e3:.2f} ms  {nbytes/dt/1e9:.1f} GB/s")
# 2. loops
def loop(n, with_scene, with_dl, tag, nap=0.0):
    h.download_wait(); torch.cuda.synchronize()
    ph = np.zeros(3)
    t0 = time.perf_counter()
    for i in range(n):
        a = time.perf_counter()
        if with_scene: h.set_scene(scene)
        b = time.perf_counter()
        mm = h.remesh(bb, init, levels, download=False)
        c = time.perf_counter()
        if with_dl: h.download_into_async(mm, *(x.data_ptr() for x in bufs[i & 1]))
        if nap: time.sleep(nap)
        e = time.perf_counter()
        ph += [b - a, c - b, e - c]
    h.download_wait()
    tot = time.perf_counter() - t0
    print(f"{tag}: {tot/n*1e3:.2f} ms/step; host phases (ms) set_scene {ph[0]/n*1e3:.2f} remesh {ph[1]/n*1e3:.2f} dl_issue {ph[2]/n*1e3:.2f}; last_gpu_ms {h.stats()['last_gpu_ms']:.2f}")
loop(10, False, False, "remesh only")
loop(10, True, False, "set_scene + remesh")
loop(10, False, True, "remesh + async download")
loop(10, False, True, "remesh + async download + 0.3 ms nap", nap=0.0003)
loop(10, True, True, "full e2e")
loop(10, True, True, "full e2e (again)")
