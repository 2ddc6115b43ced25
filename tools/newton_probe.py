#!/usr/bin/env python3
"""Developer probe: Newton iteration histogram of the real edge mid-points of a workload (GPU)."""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import bsdmg_b200
from bsdmg_b200 import scenes

name, init, levels = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
scene = scenes.many_primitives(int(name[4:])) if name.startswith("many") else scenes.SCENES[name]()
h = bsdmg_b200.CudaHandler(0, scene)
h.field_reset(5.0, init)
for _ in range(levels):
    n = h.field_refine()
print("voxels", n)
vox = h.field_download()
K = min(len(vox), int(sys.argv[4]) if len(sys.argv) > 4 else 20000)
_, vs = h.field_count()
sub = bsdmg_b200.CudaVoxelField(vox[:: max(1, len(vox) // K)][:K].copy(), vs)
t = time.time(); m = h.voxel_field_to_mesh(sub); print("mesh of subset", time.time() - t, "s", m.triangle_count, "tris")
st = h.stats()
U = st["unique_vertices"]
start = h.debug_fetch("ustart", 3 * U).reshape(-1, 3)
t = time.time(); out, it = h.eval_project(start); print("project", time.time() - t, "s")
print("iters: mean %.2f p50 %d p99 %d max %d  n>=100: %d  n>=10000: %d of %d" % (it.mean(), np.percentile(it, 50), np.percentile(it, 99), it.max(), (it >= 100).sum(), (it >= 10000).sum(), U))
bad = np.argsort(it)[-5:]
for b in bad:
    print(" worst", it[b], start[b], "->", out[b], "sd", h.eval_sdf(out[b][None])[0], "n", h.eval_normal(out[b][None])[0])
np.save("gpurun_out/newton_start_%s.npy" % name, start[np.argsort(it)[-50:]])
