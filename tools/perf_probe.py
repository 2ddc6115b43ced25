#!/usr/bin/env python3
"""Developer probe: per-kernel CUDA-event times of one remesh per workload (not the bench)."""
import sys, time, pathlib, os, ctypes
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import bsdmg_b200
from bsdmg_b200 import scenes

def run(name, scene, bb, init, levels, reps=3):
    h = bsdmg_b200.CudaHandler(0, scene)
    h.set_profiling(True)
    for r in range(reps):
        if r == reps - 1 and os.environ.get("SDM_PROFILE_LAST"):   # ncu --profile-from-start off: only the last remesh is captured
            ctypes.CDLL("libcudart.so").cudaProfilerStart()
        t = time.time(); m = h.remesh(bb, init, levels, download=False); wall = time.time() - t
    st = h.stats()
    print(f"== {name}: res {init << levels}^3 voxels {st['level_counts'][:levels+1]} tris {m.triangle_count} verts {m.vertex_count} uniq {st['unique_vertices']} "
          f"gpu {st['last_gpu_ms']:.3f} ms wall {wall*1e3:.3f} ms evals {st['sdf_evals']/1e6:.1f}M -> {st['sdf_evals']/st['last_gpu_ms']/1e6:.2f} Gevals/s")
    pe = st["prim_evals"]; lc = st["level_counts"]
    ev = dict(refine=27 * sum(lc[:levels]), classify=8 * lc[levels], normals=12 * st["unique_vertices"], orient=12 * st["raw_triangles"])
    ev["project"] = st["sdf_evals"] - sum(ev.values())
    # (orient: 12 per triangle as the algorithm states it; the six-sample test really evaluates about half of them)
    print("     primitives folded per evaluation: " + ", ".join(f"{k} {pe[k] / max(v, 1):.1f}" for k, v in ev.items()))
    for k, ms in h.kernel_times():
        print(f"     {k:18s} {ms*1e3:9.1f} us")
    h.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["sd_obj_1024", "sd_obj_512_c2"]
    for w in which:
        if w == "sd_obj_1024": run(w, scenes.sd_obj(), 5.0, 32, 5)
        elif w == "sd_obj_512_c2": run(w, scenes.sd_obj(), 5.0, 64, 3)
        elif w == "sd_obj_1024_i64": run(w, scenes.sd_obj(), 5.0, 64, 4)
        elif w == "many1024_256": run(w, scenes.many_primitives(1024), 5.0, 64, 2)
        elif w == "many1024_1024": run(w, scenes.many_primitives(1024), 5.0, 64, 4)
        elif w == "many64_512": run(w, scenes.many_primitives(64), 5.0, 64, 3)
        elif w == "mandelbulb_512": run(w, scenes.mandelbulb(), 5.0, 32, 4)
        elif w == "sphere_box_128": run(w, scenes.sphere_box(), 5.0, 32, 2)
