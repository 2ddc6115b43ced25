#!/usr/bin/env python3
"""Developer probe: single frames of the animated 1 024-primitive scene (configs[4]) with per-kernel times and the tail kernel's
debug counters (warp steps, list rebuilds, steps without a tile list).  usage: c5_frame_probe.py frame [frame ...]"""
import sys, struct, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import bsdmg_b200
from bsdmg_b200 import scenes

frames = [int(a) for a in sys.argv[1:]] or [254, 329]
h = bsdmg_b200.CudaHandler(0, scenes.many_primitives(1024, t=0.0))
h.remesh(5.0, 64, 4, download=False)
h.set_profiling(True)
for f in frames:
    h.set_scene(scenes.many_primitives(1024, t=f / 60.0))
    for rep in range(2):
        h.remesh(5.0, 64, 4, download=False)
        st = h.stats()
        kt = {}
        for name, ms in h.kernel_times():
            kt[name] = kt.get(name, 0.0) + ms
        raw = h.debug_fetch("state", 76, np.uint32).tobytes()   # sizeof(DevState) = 55 u32 (+ pad) + 10 u64 = 304
        tail = struct.unpack_from("<3Q", raw, len(raw) - 24)
        pe = struct.unpack_from("<6Q", raw, len(raw) - 24 - 8 - 48)
        print(f"frame {f} rep {rep}: gpu {st['last_gpu_ms']:.2f} ms tail {kt.get('k_project_tail', 0):.2f} ms stragglers {st['stragglers']} "
              f"tail steps/rebuilds/unlisted {tail} tail prim evals {pe[3]} -> mean list {pe[3] / max(1, 13 * tail[0]):.1f} (per running half)")
