#!/usr/bin/env python3
"""Developer probe: how many primitives survive in the per-cell masks of the 1024-primitive scene."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import bsdmg_b200
from bsdmg_b200 import scenes
init = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h = bsdmg_b200.CudaHandler(0, scenes.many_primitives(1024))
h.field_reset(5.0, init)
n = h.field_refine()
G = init
W = 32
m = h.debug_fetch("masks_fine", G**3 * W, np.uint32).reshape(G, G, G, W)
pc = np.unpackbits(m.view(np.uint8), axis=-1).reshape(G, G, G, -1).sum(-1)
print("all cells: mean K %.1f median %d p90 %d max %d" % (pc.mean(), np.median(pc), np.percentile(pc, 90), pc.max()))
h.field_reset(5.0, init)
h.field_refine()
v1 = h.field_download()
cell = 5.0 / G
idx = np.floor((v1 + 2.5) / cell + 1e-4).astype(int).clip(0, G - 1)
act = np.zeros((G, G, G), bool); act[idx[:, 0], idx[:, 1], idx[:, 2]] = True
print("active cells %d: mean K %.1f median %d p90 %d max %d" % (act.sum(), pc[act].mean(), np.median(pc[act]), np.percentile(pc[act], 90), pc[act].max()))
Gc = G // 4
mc = h.debug_fetch("masks_coarse", Gc**3 * W, np.uint32).reshape(Gc, Gc, Gc, W)
pcc = np.unpackbits(mc.view(np.uint8), axis=-1).reshape(Gc, Gc, Gc, -1).sum(-1)
print("coarse cells: mean K %.1f max %d" % (pcc.mean(), pcc.max()))
