"""Host-side logic of the multi-GPU path on CPU: shard arithmetic, offset planning, and - over a real 2-process gloo
group - that concatenating per-shard triangle soups in rank order and welding once gives exactly the single-process
mesh (the property the NCCL gather relies on; the shards here are produced by the CPU oracle)."""
import os
import socket

import numpy as np
import pytest

from bsdmg_b200 import parallel, scenes


def test_shard_ranges_tile_the_list():
    for n in (0, 1, 7, 4136, 1062512):
        for count in (1, 2, 3, 8):
            ranges = [parallel.shard_range(n, s, count) for s in range(count)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_plan_offsets():
    v, t, tot = parallel.plan_offsets([(5, 7), (0, 0), (3, 2)])
    assert v == [0, 5, 5] and t == [0, 7, 7] and tot == (8, 9)


def test_split_level_choice():
    assert parallel.choose_split_level(32, 5, 1) == 0
    assert parallel.choose_split_level(32, 5, 8) == 2
    assert parallel.choose_split_level(128, 4, 8) == 1
    assert parallel.choose_split_level(64, 0, 4) == 0
    assert parallel.choose_split_level(64, 4, 8, culled=True) == 0     # culled scenes: level-0 split by the cells' surface flags
    assert parallel.scene_is_culled(scenes.many_primitives(1024)) and not parallel.scene_is_culled(scenes.sd_obj())
    assert not parallel.scene_is_culled(scenes.mandelbulb()) and not parallel.scene_is_culled(scenes.many_primitives(16))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    scene = scenes.sd_obj()
    o = orc.Oracle(scene)
    orc.Oracle.set_threads(2)
    vox, vs = o.create_voxel_field(5.0, 32)
    vox, vs = o.refine(vox, vs)                      # split level 1: every rank redundantly
    lo, hi = parallel.shard_range(vox.shape[0], rank, world)
    mine, mvs = o.refine(vox[lo:hi], vs)             # the shard, refined on its own
    tris, _ = o.mesh_raw(mine, mvs)
    tris = tris[np.isfinite(tris[:, 0])]             # the shard's triangles (finite filter, src/cuda/mod.rs:289)
    counts = torch.zeros(world, dtype=torch.int64)
    mine_n = torch.tensor([tris.shape[0]], dtype=torch.int64)
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine_n)
    counts = [int(g[0]) for g in gathered]
    _, t_off, (_, T) = parallel.plan_offsets([(0, c) for c in counts])
    if rank == 0:
        allt = np.empty((T, 18), np.float32)
        allt[: counts[0]] = tris
        for r in range(1, world):
            buf = torch.empty((counts[r], 18), dtype=torch.float32)
            dist.recv(buf, src=r)
            allt[t_off[r]: t_off[r] + counts[r]] = buf.numpy()
        pos, nrm, idx = orc.Oracle.weld(allt)
        full = o.remesh(5.0, 32, 2)
        ok = (np.array_equal(idx, full["indices"]) and np.array_equal(pos.view(np.uint32), full["positions"].view(np.uint32))
              and np.array_equal(nrm.view(np.uint32), full["normals"].view(np.uint32)))
        q.put(bool(ok))
    else:
        dist.send(torch.from_numpy(np.ascontiguousarray(tris)), dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_reproduces_single_mesh():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_boundary_intervals():
    ranges = [(-2.0, -0.5), (-0.6, 0.7), (1.0, 0.0), (0.6, 2.0)]          # shard 2 has no finite vertex (min > max)
    iv = parallel.boundary_intervals(ranges, 1)
    assert len(iv) == 2 and iv[0] == (-2.0 - 1e-4, -0.5 + 1e-4) and iv[1] == (0.6 - 1e-4, 2.0 + 1e-4)
    assert parallel.boundary_intervals([(0.0, 1.0)], 0) == []


def test_weld_keys_follow_the_reference():
    p = np.array([[0.0, 1.234565, -1.234565], [np.nan, 2.5e-6, -2.5e-6], [1e30, -1e30, 0.000005]], np.float32)
    k = parallel.weld_keys(p)
    f = lambda x: float(np.float32(x) * np.float32(10e4))
    assert k[0].tolist() == [0, int(np.floor(abs(f(1.234565)) + 0.5)), -int(np.floor(abs(f(1.234565)) + 0.5))]
    assert k[1, 0] == 0 and k[2, 0] == np.iinfo(np.int64).max and k[2, 1] == np.iinfo(np.int64).min


def _dw_worker(rank, world, port, q):
    """Distributed weld over a real gloo group: every rank welds its own shard, rank 0 resolves the shared keys from the
    boundary candidates only and concatenates - the result must be the single-process mesh, byte for byte."""
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    o = orc.Oracle(scenes.sd_obj())
    orc.Oracle.set_threads(2)
    vox, vs = o.create_voxel_field(5.0, 32)
    vox, vs = o.refine(vox, vs)
    lo, hi = parallel.shard_range(vox.shape[0], rank, world)
    mine, mvs = o.refine(vox[lo:hi], vs)
    tris, _ = o.mesh_raw(mine, mvs)
    pos, nrm, idx = orc.Oracle.weld(tris)                                   # local weld (the finite filter is part of it)
    rng = torch.tensor([float(pos[:, 0].min()), float(pos[:, 0].max()), pos.shape[0], idx.shape[0]], dtype=torch.float64)
    allr = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allr, rng)
    ranges = [(float(r[0]), float(r[1])) for r in allr]
    x = pos[:, 0]
    cand = np.zeros(pos.shape[0], dtype=bool)
    for a, b in parallel.boundary_intervals(ranges, rank):
        cand |= (x >= a) & (x <= b)
    if rank == 0:
        shards, masks = [(pos, nrm, idx)], [cand]
        for r in range(1, world):
            V, T = int(allr[r][2]), int(allr[r][3])
            bp, bn, bi, bm = (torch.empty((V, 3), dtype=torch.float32), torch.empty((V, 3), dtype=torch.float32),
                              torch.empty((T, 3), dtype=torch.int32), torch.empty(V, dtype=torch.uint8))
            for b in (bp, bn, bi, bm):
                dist.recv(b, src=r)
            shards.append((bp.numpy(), bn.numpy(), bi.numpy().view(np.uint32)))
            masks.append(bm.numpy().astype(bool))
        full = o.remesh(5.0, 32, 2)
        ok = True
        for m in (masks, None):                                              # candidates only == all vertices
            p2, n2, i2 = parallel.concat_welded_shards(shards, m)
            ok = ok and (np.array_equal(i2, full["indices"]) and np.array_equal(p2.view(np.uint32), full["positions"].view(np.uint32))
                         and np.array_equal(n2.view(np.uint32), full["normals"].view(np.uint32)))
        shared = int(sum(m.sum() for m in masks))
        q.put((bool(ok), shared, int(sum(s[0].shape[0] for s in shards)) - int(full["positions"].shape[0])))
    else:
        for a in (pos, nrm, idx.view(np.int32), cand.astype(np.uint8)):
            dist.send(torch.from_numpy(np.ascontiguousarray(a)), dst=0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_distributed_weld_reproduces_single_mesh(world):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dw_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, candidates, removed = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert removed > 0 and candidates >= 2 * removed      # the shards do share vertices along their interfaces
