"""Shard / merge parity on ONE GPU ("fake multi-GPU": G handles on the same device, device-to-device copies instead
of NCCL): the merged mesh must be byte-identical to the single-handle mesh, for even and ragged shard counts."""
import numpy as np
import pytest

import bsdmg_b200
from bsdmg_b200 import parallel, scenes

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("scene_name,init,levels,split,G", [("sd_obj", 32, 3, 1, 2), ("sd_obj", 32, 3, 2, 3), ("sd_obj", 32, 2, 0, 4),
                                                            ("sphere_box", 32, 2, 1, 2), ("sd_obj", 32, 2, 2, 7)])
def test_sharded_mesh_equals_single(scene_name, init, levels, split, G):
    import torch

    scene = scenes.SCENES[scene_name]()
    dev = torch.device("cuda", 0)
    hs = [bsdmg_b200.CudaHandler(0, scene) for _ in range(G)]
    try:
        single = hs[0].remesh(5.0, init, levels)
        infos = [h.shard_remesh(5.0, init, levels, split, r, G) for r, h in enumerate(hs)]
        # the shards tile the split-level list exactly
        assert infos[0]["voxel_begin"] == 0 and infos[-1]["voxel_end"] == infos[0]["split_total"]
        for a, b in zip(infos[:-1], infos[1:]):
            assert a["voxel_end"] == b["voxel_begin"]
        for r, info in enumerate(infos):
            assert (info["voxel_begin"], info["voxel_end"]) == parallel.shard_range(info["split_total"], r, G)
        counts = [(i["unique_vertices"], i["raw_triangles"]) for i in infos]
        v_off, t_off, (V, T) = parallel.plan_offsets(counts)
        hs[0].shard_reserve(V, T)
        root = hs[0].shard_buffers()
        for r in range(1, G):
            hs[r].shard_prepare_send(v_off[r])
            b = hs[r].shard_buffers()
            u, tr = counts[r]
            for key, n, off, ts in (("positions", 3 * u, 12 * v_off[r], "<f4"), ("normals", 3 * u, 12 * v_off[r], "<f4"),
                                    ("triangle_vertex_ids", 3 * tr, 12 * t_off[r], "<i4")):
                if n:
                    parallel._view(torch, root[key] + off, n, ts, dev).copy_(parallel._view(torch, b[key], n, ts, dev))
        torch.cuda.synchronize()
        merged = hs[0].shard_weld(V, T, download=True)
        assert merged.triangle_count == single.triangle_count and merged.vertex_count == single.vertex_count
        assert np.array_equal(merged.indices, single.indices)
        assert np.array_equal(bits(merged.positions), bits(single.positions))
        assert np.array_equal(bits(merged.normals), bits(single.normals))
    finally:
        for h in hs:
            h.close()
