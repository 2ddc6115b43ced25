#!/usr/bin/env python3
"""Developer probe: the vertices a frame of the animated scene (configs[4]) hands to the tail kernel - start point, state at the
hand-over, final position - saved for a CPU trace of their Newton orbits.  usage: c5_straggler_dump.py frame"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import bsdmg_b200
from bsdmg_b200 import scenes

f = int(sys.argv[1]) if len(sys.argv) > 1 else 254
h = bsdmg_b200.CudaHandler(0, scenes.many_primitives(1024, t=f / 60.0))
h.remesh(5.0, 64, 4, download=False)
st = h.stats()
n, U = st["stragglers"], st["unique_vertices"]
rec = h.debug_fetch("stragglers", 12 * n, np.uint32).reshape(n, 12)
start = h.debug_fetch("ustart", 3 * U).reshape(-1, 3)
pos = h.debug_fetch("upos", 3 * U).reshape(-1, 3)
uid = rec[:, 0]
out = np.concatenate([uid[:, None].astype(np.float64), rec[:, 1:2].astype(np.float64), start[uid].astype(np.float64),
                      rec[:, 2:5].copy().view(np.float32).astype(np.float64), pos[uid].astype(np.float64)], axis=1)
np.save(f"gpurun_out/stragglers_frame{f}.npy", out)
np.save(f"gpurun_out/stragglers_frame{f}_startbits.npy", start[uid].view(np.uint32))
for r in out:
    print("uid %d it %d start %s handed %s final %s" % (r[0], r[1], r[2:5], r[5:8], r[8:11]))
