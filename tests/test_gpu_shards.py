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


@pytest.mark.parametrize("scene_name,init,levels,split,G", [("sd_obj", 32, 3, 1, 2), ("sd_obj", 32, 3, 2, 3), ("many64", 32, 2, 1, 4),
                                                            ("sphere_box", 32, 2, 1, 2), ("sd_obj", 32, 2, 2, 7), ("sd_obj", 32, 2, 0, 1)])
def test_distributed_weld_equals_single(scene_name, init, levels, split, G):
    """The preferred exchange (include/sdfmesh.h, "Distributed weld"): every shard welds locally, rank 0 resolves the keys the
    shards share along their interfaces, duplicates are dropped and re-mapped, the shards are concatenated.  Same bytes as the
    single-handle mesh."""
    import torch

    scene = scenes.many_primitives(64) if scene_name == "many64" else scenes.SCENES[scene_name]()
    dev = torch.device("cuda", 0)
    hs = [bsdmg_b200.CudaHandler(0, scene) for _ in range(G)]
    copy = lambda dst, src, n, ts: n and parallel._view(torch, dst, n, ts, dev).copy_(parallel._view(torch, src, n, ts, dev))
    try:
        single = hs[0].remesh(5.0, init, levels)
        for r, h in enumerate(hs):
            h.shard_remesh(5.0, init, levels, split, r, G)
        ws = [h.shard_local_weld() for h in hs]
        assert all(w["nonfinite"] == 0 for w in ws)
        rows = []
        for r, h in enumerate(hs):
            ivals = [(w["min_x"] - 1e-4, w["max_x"] + 1e-4) for q, w in enumerate(ws) if q != r and w["min_x"] <= w["max_x"]]
            rows.append(h.shard_boundary_keys(ivals))
        v_off, t_off, (VS, T) = parallel.plan_offsets([(w["vertices"], w["triangles"]) for w in ws])
        # (rank 0 reserves before it extracts its own key rows: a re-allocation would lose them - so it extracts them again here)
        hs[0].shard_reserve_welded(VS, T)
        n0 = rows[0][1]
        rows[0] = hs[0].shard_boundary_keys([(w["min_x"] - 1e-4, w["max_x"] + 1e-4) for q, w in enumerate(ws) if q != 0 and w["min_x"] <= w["max_x"]])
        assert rows[0][1] == n0
        k_off, _, (K, _) = parallel.plan_offsets([(n, 0) for _, n in rows])
        base = hs[0].shard_key_scratch(K)
        root = hs[0].shard_welded_buffers()
        for r in range(1, G):
            copy(base + 16 * k_off[r], rows[r][0], 4 * rows[r][1], "<i4")
            b = hs[r].shard_welded_buffers()
            copy(root["positions"] + 12 * v_off[r], b["positions"], 3 * ws[r]["vertices"], "<f4")
            copy(root["normals"] + 12 * v_off[r], b["normals"], 3 * ws[r]["vertices"], "<f4")
            copy(root["indices"] + 12 * t_off[r], b["indices"], 3 * ws[r]["triangles"], "<i4")
        torch.cuda.synchronize()
        removed = hs[0].shard_resolve(base, K, [w["vertices"] for w in ws])
        assert removed[0] == 0
        if G > 1 and scene_name == "sd_obj":
            assert sum(removed) > 0                              # neighbouring shards mesh the edges of their interface twice
        merged = hs[0].shard_fixup([w["triangles"] for w in ws], download=True)
        assert (merged.vertex_count, merged.triangle_count) == (VS - sum(removed), T) == (single.vertex_count, single.triangle_count)
        assert np.array_equal(merged.indices, single.indices)
        assert np.array_equal(bits(merged.positions), bits(single.positions))
        assert np.array_equal(bits(merged.normals), bits(single.normals))
    finally:
        for h in hs:
            h.close()


@pytest.mark.parametrize("scene_name,init,levels,split,G", [("sd_obj", 32, 3, 1, 2), ("sd_obj", 32, 3, 2, 3), ("many64", 32, 2, 1, 4),
                                                            ("many256", 32, 3, 1, 4), ("sphere_box", 32, 2, 1, 2), ("sd_obj", 32, 2, 2, 7),
                                                            ("sd_obj", 32, 2, 0, 1)])
@pytest.mark.parametrize("deliver", [0, 1])
def test_peer_exchange_equals_single(scene_name, init, levels, split, G, deliver):
    """The device-driven exchange (include/sdfmesh.h, "peer exchange") with the ranks emulated by G handles on one GPU (phases
    issued one after the other: no kernel waits for another one here).  deliver = 0: the merged mesh in rank 0's second output
    set; deliver = 1: every rank's own rows, assembled on the host at the offsets the step reports.  Same bytes as the
    single-handle mesh either way, over two consecutive steps (the payload slots alternate with the epoch)."""
    scene = scenes.many_primitives(int(scene_name[4:])) if scene_name.startswith("many") else scenes.SCENES[scene_name]()
    hs = [bsdmg_b200.CudaHandler(0, scene) for _ in range(G)]
    try:
        single = hs[0].remesh(5.0, init, levels)
        cap_vox, cap_rows = parallel.peer_capacities(single.vertex_count, single.triangle_count, G)
        hs[0].reserve(max(cap_vox, init ** 3))
        blob = hs[0].peer_root_export(G, cap_rows)
        for r, h in enumerate(hs):
            h.peer_attach(blob, r, G, same_process_root=hs[0])
        for epoch in (1, 2, 3):     # epoch 2 on: shard bounds re-balanced from the measured compute times of the step before
            out = parallel.emulated_peer_step(hs, 5.0, init, levels, split, epoch, deliver)
            res0 = out[0][0]
            assert res0["status"] == 0
            assert (res0["total_vertices"], res0["total_triangles"]) == (single.vertex_count, single.triangle_count)
            if deliver == 0:
                merged = hs[0]._download(out[0][1])
            else:
                pos = np.empty((single.vertex_count, 3), np.float32); nrm = np.empty_like(pos)
                idx = np.empty((single.triangle_count, 3), np.uint32)
                v_end = t_end = 0
                for h, (res, m) in zip(hs, out):
                    part = h._download(m)
                    assert part.vertex_count == res["vertices"] and part.triangle_count == res["triangles"]
                    assert res["vertex_offset"] == v_end and res["triangle_offset"] == t_end      # the ranks' rows tile the mesh in rank order
                    pos[v_end:v_end + res["vertices"]] = part.positions; nrm[v_end:v_end + res["vertices"]] = part.normals
                    idx[t_end:t_end + res["triangles"]] = part.indices
                    v_end += res["vertices"]; t_end += res["triangles"]
                assert (v_end, t_end) == (single.vertex_count, single.triangle_count)
                merged = bsdmg_b200.Mesh(pos, nrm, idx)
            assert np.array_equal(merged.indices, single.indices)
            assert np.array_equal(bits(merged.positions), bits(single.positions))
            assert np.array_equal(bits(merged.normals), bits(single.normals))
    finally:
        for h in hs:
            h.close()
