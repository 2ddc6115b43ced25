#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (one tool per call): sd_obj 32^3 -> 64^3, a culled 64-primitive scene,
a 2-shard merge and the host-buffer API."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import bsdmg_b200
from bsdmg_b200 import scenes
h = bsdmg_b200.CudaHandler(0)
m = h.remesh(5.0, 32, 1)
print("sd_obj", m.triangle_count)
h.set_scene(scenes.many_primitives(64))
m = h.remesh(5.0, 16, 2)
print("many64", m.triangle_count)
f = bsdmg_b200.CudaHandler.create_cuda_voxel_field(5.0, 16)
h.refine_voxel_field(f); m = h.voxel_field_to_mesh(f)
print("host api", len(f), m.triangle_count)
h.set_scene(scenes.sd_obj())
info = h.shard_remesh(5.0, 32, 1, 1, 0, 2)
print("shard", info["unique_vertices"], info["raw_triangles"])
m = h.shard_weld(info["unique_vertices"], info["raw_triangles"], download=True)
print("shard weld", m.triangle_count)
h.close()
