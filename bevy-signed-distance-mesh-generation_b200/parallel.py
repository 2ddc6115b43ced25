"""Multi-GPU remesh: contiguous shards of the active list, one process / handle per GPU; every shard is welded where it was
made, the keys shared along the shard interfaces are resolved on rank 0, and the welded shards are gathered to rank 0 over
NVLink with NCCL (torch.distributed) into exactly the single-GPU mesh.

Why this shape: every voxel's classification, vertices and normals depend only on its own corner
coordinates and the analytic SDF (compute_mesh_generation.cu:27-58, 74-86) - no halo is needed - and the list is
x-major with children in parent order, so a contiguous part of the list is an x-slab at every level.  Only the weld
(global first-occurrence order, src/cuda/mod.rs:263-296) needs all shards: it is the one exchange step.

`plan_offsets` and `merge_counts` are pure host logic (tested with gloo on CPU, world size 2); the transport uses
zero-copy torch views of the library's device buffers.
"""
from __future__ import annotations

import time

import numpy as np


def shard_range(n: int, shard: int, count: int):
    """The contiguous part of an n-voxel list owned by `shard` (same arithmetic as k_take_shard)."""
    return (n * shard) // count, (n * (shard + 1)) // count


def plan_offsets(counts):
    """counts: [(unique_vertices, raw_triangles)] per rank, in rank order -> (vertex_offsets, triangle_offsets, totals)."""
    v_off, t_off = [], []
    v = t = 0
    for u, tr in counts:
        v_off.append(v)
        t_off.append(t)
        v += int(u)
        t += int(tr)
    return v_off, t_off, (v, t)


def boundary_intervals(ranges, rank: int, widen: float = 1e-4):
    """x intervals in which a welded vertex of shard `rank` can share its quantised key (src/cuda/mod.rs:270: round(x * 1e5)) with
    a vertex of another shard: the other shards' [min_x, max_x], widened.  Two equal keys are less than 1.1e-5 apart, so a
    vertex outside all of these intervals is unique to its shard.  Shards without finite vertices (min > max) contribute nothing."""
    return [(lo - widen, hi + widen) for r, (lo, hi) in enumerate(ranges) if r != rank and lo <= hi]


def weld_keys(positions: np.ndarray) -> np.ndarray:
    """The reference's weld key (src/cuda/mod.rs:270): (x * 10e4f32).round() as i64 per component, as int64 rows.
    (Rust's round is half-away-from-zero; `as i64` saturates and sends NaN to 0.)"""
    c = positions.astype(np.float32) * np.float32(10e4)
    r = np.where(np.isnan(c), np.float32(0), np.copysign(np.floor(np.abs(c) + np.float32(0.5)), c))
    r = r.astype(np.float64)
    big = 9223372036854775808.0                   # 2^63: `as i64` saturates
    out = np.where(np.abs(r) < big, r, 0.0).astype(np.int64)
    out[r >= big] = np.iinfo(np.int64).max
    out[r <= -big] = np.iinfo(np.int64).min
    return out


def concat_welded_shards(shards, candidate_masks=None):
    """Host restatement of the distributed weld (include/sdfmesh.h; sdm_shard_resolve + sdm_shard_fixup do this on rank 0's
    GPU).  shards: [(positions[V,3], normals[V,3], indices[T,3])] - each shard welded on its own, local indices, in shard
    order.  A key's owner is its copy in the LOWEST shard; other copies are removed (stable) and the indices that used them
    point to the owner.  candidate_masks (optional, per shard, bool[V]): only these vertices are looked at - what
    boundary_intervals() selects; must give the same result as looking at all of them."""
    voff = np.cumsum([0] + [p.shape[0] for p, _, _ in shards])
    owner = {}                                   # key -> concatenated position of its first copy (shards in order = lowest shard)
    removed = np.zeros(voff[-1], dtype=bool)
    remap = np.zeros(voff[-1], dtype=np.int64)
    for s, (pos, _, _) in enumerate(shards):
        keys = weld_keys(pos)
        sel = np.arange(pos.shape[0]) if candidate_masks is None else np.nonzero(candidate_masks[s])[0]
        for i in sel:
            k = tuple(int(v) for v in keys[i])
            c = int(voff[s] + i)
            if k in owner:
                removed[c] = True
                remap[c] = owner[k]
            else:
                owner[k] = c
    before = np.cumsum(removed) - removed        # removed vertices before each concatenated position
    gid = np.arange(voff[-1]) - before           # global id of a kept vertex
    final = np.where(removed, gid[remap], gid)   # an owner is never removed
    positions = np.concatenate([p for p, _, _ in shards])[~removed]
    normals = np.concatenate([n for _, n, _ in shards])[~removed]
    indices = np.concatenate([final[idx.astype(np.int64) + voff[s]] for s, (_, _, idx) in enumerate(shards)]).astype(np.uint32)
    return positions, normals, indices


def scene_is_culled(scene) -> bool:
    """The library's rule for using per-cell primitive masks (csrc/sdfmesh.cu, sdm_set_scene): more than 24 compiled primitives
    (a box skeleton counts 12) and no Mandelbulb."""
    from .scenes import BOX_SKELETON, MANDELBULB

    kinds = [int(k) for k in scene["kind"]]
    return MANDELBULB not in kinds and sum(12 if k == BOX_SKELETON else 1 for k in kinds) > 24


def choose_split_level(init_factor: int, levels: int, world: int, culled: bool = False) -> int:
    """Culled scenes split the dense level-0 list itself, balanced by the mask cells' surface flags (k_shard_bounds_by_flags):
    nothing is refined redundantly.  Otherwise coarse levels are refined on every rank and the split happens once the list is
    long enough that an equal-count split is also a balanced split of the final work, never later than level 2."""
    if world <= 1 or culled:
        return 0
    return min(levels, 1 if init_factor >= 64 else 2)


class _DevView:
    """Zero-copy view of library-owned device memory for torch (``torch.as_tensor(view, device='cuda')``)."""

    def __init__(self, ptr: int, nelem: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _view(torch, ptr, nelem, typestr, device):
    if nelem == 0:
        return torch.empty(0, dtype=torch.float32 if typestr == "<f4" else torch.int32, device=device)
    return torch.as_tensor(_DevView(ptr, nelem, typestr), device=device)


def emulated_peer_step(handlers, bb_size, init_factor, levels, split_level, epoch, deliver=0):
    """The peer exchange with the ranks emulated by several handles of ONE process on one GPU (tests): the phases are issued one
    after the other, with a synchronisation in between instead of the device-side flags (kernels that wait for each other must not
    share a GPU).  handlers[0] is rank 0.  -> [(result, mesh view)] per rank."""
    for phase in (1, 2, 4, 8, 16):
        for h in handlers:
            h.peer_step(bb_size, init_factor, levels, split_level, epoch, deliver, phase, spin=False)
        for h in handlers:
            h.sync()
    return [h.peer_finish() for h in handlers]


def peer_capacities(total_vertices: int, total_triangles: int, world: int, max_shard_vertices: int = 0):
    """(voxel capacity of rank 0, key rows per rank) for a merged mesh of the given size: rank 0's second output set must hold the whole
    mesh (2 vertices / 3 triangles per voxel of capacity), with a quarter of head room.  A rank's interface candidates are its vertices
    inside another shard's x range: where a shard boundary cuts through a layer of level-0 cells the ranges overlap by a whole cell, which
    can be a third of a thin shard - so a slot holds a whole (even) shard and a half."""
    cap_vox = int(max(total_vertices / 2.0, total_triangles / 3.0) * 1.25) + 4096
    cap_rows = max(1 << 16, int(total_vertices / max(world, 1) * 1.5), int(max_shard_vertices * 1.25))
    return cap_vox, cap_rows


class PeerRemesher:
    """step() = one full remesh on `world` GPUs (one process each) through the device-driven peer exchange (include/sdfmesh.h):
    no host round trip, no NCCL on the data path - torch.distributed only carries the set-up (counts, the IPC handles).
    deliver = 0: the merged mesh is assembled in rank 0's HBM over NVLink; deliver = 1: every rank keeps its rows and copies them
    over its own PCIe link into a host buffer shared by the ranks (`shared_host`)."""

    def __init__(self, handler, bb_size, init_factor, levels, rank, world, dist, split_level=None, culled=False):
        import torch

        self.h, self.bb, self.init, self.levels = handler, bb_size, init_factor, levels
        self.rank, self.world, self.dist = rank, world, dist
        self.split_level = choose_split_level(init_factor, levels, world, culled) if split_level is None else split_level
        self.epoch = 0
        self.last_gpu_ms = 0.0
        self.mesh = None
        self.last = None
        self._dev = torch.device("cuda", torch.cuda.current_device())
        self.shared = None
        self.growth = 1.0
        self._setup()

    def _setup(self):
        """Sizes through the host-driven calls (once, or again with more head room after a step that did not fit), capacities,
        rank 0's export, everybody's attach."""
        import torch

        h, dist, rank, world = self.h, self.dist, self.rank, self.world
        info = h.shard_remesh(self.bb, self.init, self.levels, self.split_level, rank, world)
        w = h.shard_local_weld()
        mine = torch.tensor([w["vertices"], w["triangles"]], dtype=torch.int64, device=self._dev)
        allc = torch.empty((world, 2), dtype=torch.int64, device=self._dev)
        dist.all_gather_into_tensor(allc, mine)
        totals = allc.sum(dim=0).cpu().tolist()
        # (wandering Newton iterates can stretch a shard's x range over the whole domain - Mandelbulb - and then every vertex of the
        # other shards is an interface candidate: a slot must hold the largest shard)
        self.cap_vox, self.cap_rows = peer_capacities(int(totals[0]), int(totals[1]), world, int(allc[:, 0].max().item() * self.growth * 1.5))
        # every rank: head room for a shard that the measured load balancing makes larger than the even split's
        h.reserve(max(int(max(info["final_voxels"], w["vertices"] / 2.0, w["triangles"] / 3.0) * 1.6 * self.growth) + 4096, self.init ** 3))
        if rank == 0:
            h.reserve(int(self.cap_vox * self.growth))
        blob = [h.peer_root_export(world, self.cap_rows) if rank == 0 else None]
        dist.broadcast_object_list(blob, src=0)
        h.peer_attach(blob[0], rank, world)
        torch.cuda.synchronize()
        dist.barrier()

    def _recover(self):
        """A step failed on all ranks alike (a shard, the merged mesh or a key-row slot did not fit): unmap, grow, set up again."""
        self.h.peer_detach()
        self.dist.barrier()
        self.growth *= 1.5
        self._setup()

    def step(self, deliver=0):
        from .handler import SdfMeshError

        for attempt in range(3):
            self.epoch += 1
            self.h.peer_step(self.bb, self.init, self.levels, self.split_level, self.epoch, deliver, 31, True)
            try:
                res, m = self.h.peer_finish()
                break
            except SdfMeshError as exc:        # every rank sees the same status word, so all of them take this branch together
                if exc.code != 4 or attempt == 2:
                    raise
                self._recover()
        self.last, self.mesh = res, (m if (deliver == 0 and self.rank == 0) else None)
        self.last_gpu_ms = res["gpu_ms"]
        return {"triangles": int(res["total_triangles"]), "vertices": int(res["total_vertices"])}

    # -- end-to-end: scene from the host every step, every rank's rows into ONE host buffer over its own PCIe link ------------
    def _shared_host(self, total_vertices, total_triangles):
        """A host buffer shared by the ranks (POSIX shared memory, page-locked in every process): positions | normals | indices."""
        import torch
        from multiprocessing import shared_memory

        nv, nt = int(total_vertices * 1.25) + 1024, int(total_triangles * 1.25) + 1024
        size = nv * 24 + nt * 12
        name = [None]
        if self.rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=size)
            name[0] = shm.name
        self.dist.broadcast_object_list(name, src=0)
        if self.rank != 0:
            shm = shared_memory.SharedMemory(name=name[0])
        import ctypes

        base = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
        rc = torch.cuda.cudart().cudaHostRegister(base, size, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed: {rc}")
        self.shared = dict(shm=shm, base=base, size=size, pos=base, nrm=base + nv * 12, idx=base + nv * 24, nv=nv, nt=nt)
        self.dist.barrier()
        return self.shared

    def e2e(self, scene, steps, warmup, barrier):
        h = self.h
        scene = np.ascontiguousarray(scene)
        if self.shared is None:
            r = self.step(deliver=1)
            self._shared_host(r["vertices"], r["triangles"])
        sh = self.shared
        d2h = 0
        barrier()
        t0 = time.perf_counter()
        for i in range(warmup + steps):
            if i == warmup:
                h.download_wait()
                barrier()
                t0 = time.perf_counter()
            h.set_scene(scene)                 # host -> device: the scene table (the path's only input), on every rank
            r = self.step(deliver=1)
            if r["vertices"] > sh["nv"] or r["triangles"] > sh["nt"]:
                raise RuntimeError("shared host buffer too small")
            h.peer_download_async(self.last, sh["pos"], sh["nrm"], sh["idx"])   # device -> host: this rank's rows, over its own link
            d2h = self.last["vertices"] * 24 + self.last["triangles"] * 12
        h.download_wait()
        barrier()
        return {"elapsed": time.perf_counter() - t0, "h2d": int(scene.nbytes), "d2h": d2h}

    def close(self):
        if self.shared is not None:
            import torch

            torch.cuda.cudart().cudaHostUnregister(self.shared["base"])
            shm = self.shared["shm"]
            self.shared = None
            try:
                shm.close()
                if self.rank == 0:
                    shm.unlink()
            except Exception:
                pass


class ShardedRemesher:
    """step() = one full remesh on `world` GPUs; on rank 0 the welded mesh is left in HBM."""

    def __init__(self, handler, bb_size, init_factor, levels, rank=0, world=1, dist=None, culled=False):
        self.h, self.bb, self.init, self.levels = handler, bb_size, init_factor, levels
        self.rank, self.world, self.dist = rank, world, dist
        self.split_level = choose_split_level(init_factor, levels, world, culled)
        self.last_gpu_ms = 0.0
        self.mesh = None
        self._pinned = None

    # -- single GPU: the fused device-resident path -------------------------------------------------
    def _step_single(self):
        m = self.h.remesh(self.bb, self.init, self.levels, download=False)
        self.last_gpu_ms = self.h.stats()["last_gpu_ms"]
        self.mesh = m
        return {"triangles": int(m.triangle_count), "vertices": int(m.vertex_count)}

    # -- N GPUs ---------------------------------------------------------------------------------------
    def _exchange(self, torch, dev, send, recv_root):
        """One grouped NCCL send/recv towards rank 0.  send: [(device pointer, element count, typestr)] of this rank (ranks > 0);
        recv_root(r): the same for rank r's rows inside rank 0's buffers."""
        dist = self.dist
        ops = []
        if self.rank == 0:
            for r in range(1, self.world):
                ops += [dist.P2POp(dist.irecv, _view(torch, p, n, ts, dev), r) for p, n, ts in recv_root(r) if n]
        else:
            ops += [dist.P2POp(dist.isend, _view(torch, p, n, ts, dev), 0) for p, n, ts in send if n]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        torch.cuda.current_stream().synchronize()

    def _step_sharded(self):
        """Local shard -> local weld -> the keys shared along the shard interfaces are resolved on rank 0 -> every rank drops its
        duplicates and makes its indices global -> the welded shards are concatenated on rank 0 (include/sdfmesh.h, "Distributed
        weld").  If that is not possible (non-finite vertices, scratch too small) the intermediate lists are gathered and rank 0
        welds everything, as in the first version."""
        import torch

        dist, h = self.dist, self.h
        dev = torch.device("cuda", torch.cuda.current_device())
        tm = [time.perf_counter()]
        info = h.shard_remesh(self.bb, self.init, self.levels, self.split_level, self.rank, self.world)
        gpu_ms = h.stats()["last_gpu_ms"]
        w = h.shard_local_weld()
        gpu_ms += h.stats()["last_gpu_ms"]
        tm.append(time.perf_counter())
        mine = torch.tensor([w["vertices"], w["triangles"], w["nonfinite"], info["unique_vertices"], info["raw_triangles"], w["min_x"], w["max_x"]],
                            dtype=torch.float64, device=dev)
        allc = torch.empty((self.world, 7), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allc, mine)
        rows_all = allc.cpu().tolist()
        counts = [[int(x) for x in row[:5]] for row in rows_all]
        ranges = [(row[5], row[6]) for row in rows_all]
        fallback = any(c[2] for c in counts) or self.world > 32
        out = {"triangles": 0, "vertices": 0}
        if not fallback:
            v_off, t_off, (VS, T) = plan_offsets([(c[0], c[1]) for c in counts])
            if self.rank == 0:
                h.shard_reserve_welded(VS, T)          # before the key rows are extracted: a re-allocation would lose them
            # candidates: my welded vertices inside another shard's x range (equal keys are < 1.1e-5 apart; widened by 1e-4)
            ivals = boundary_intervals(ranges, self.rank)
            kptr, kcnt = h.shard_boundary_keys(ivals)
            allk = torch.empty((self.world,), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allk, torch.tensor([kcnt], dtype=torch.int64, device=dev))
            kcounts = [int(x) for x in allk.cpu().tolist()]
            k_off = plan_offsets([(k, 0) for k in kcounts])[0]
            tm.append(time.perf_counter())
            b = h.shard_welded_buffers()
            if self.rank == 0:
                base = h.shard_key_scratch(sum(kcounts))   # rank 0's own rows already sit at the front of this scratch
                self._exchange(torch, dev, None, lambda r: [(base + 16 * k_off[r], 4 * kcounts[r], "<i4")])
                # the welded shards (local indices, duplicates included) arrive at their concatenated offsets while the keys are resolved
                ops = []
                for r in range(1, self.world):
                    ops += [dist.P2POp(dist.irecv, _view(torch, p, n, ts, dev), r) for p, n, ts in
                            ((b["positions"] + 12 * v_off[r], 3 * counts[r][0], "<f4"), (b["normals"] + 12 * v_off[r], 3 * counts[r][0], "<f4"),
                             (b["indices"] + 12 * t_off[r], 3 * counts[r][1], "<i4")) if n]
                works = dist.batch_isend_irecv(ops) if ops else []
                removed = h.shard_resolve(base, sum(kcounts), [c[0] for c in counts])
                tm.append(time.perf_counter())
                for q in works:
                    q.wait()
                torch.cuda.current_stream().synchronize()
                m = h.shard_fixup([c[1] for c in counts])
                gpu_ms += h.stats()["last_gpu_ms"]
                self.mesh = m
                out = {"triangles": int(m.triangle_count), "vertices": int(m.vertex_count)}
            else:
                self._exchange(torch, dev, [(kptr, 4 * kcnt, "<i4")], None)
                tm.append(time.perf_counter())
                self._exchange(torch, dev, [(b["positions"], 3 * w["vertices"], "<f4"), (b["normals"], 3 * w["vertices"], "<f4"),
                                            (b["indices"], 3 * w["triangles"], "<i4")], None)
        else:
            v_off, t_off, (V, T) = plan_offsets([(c[3], c[4]) for c in counts])
            if self.rank == 0:
                h.shard_reserve(V, T)
                b = h.shard_buffers()
                self._exchange(torch, dev, None, lambda r: [(b["positions"] + 12 * v_off[r], 3 * counts[r][3], "<f4"),
                                                            (b["normals"] + 12 * v_off[r], 3 * counts[r][3], "<f4"),
                                                            (b["triangle_vertex_ids"] + 12 * t_off[r], 3 * counts[r][4], "<i4")])
                m = h.shard_weld(V, T)
                gpu_ms += h.stats()["last_gpu_ms"]
                self.mesh = m
                out = {"triangles": int(m.triangle_count), "vertices": int(m.vertex_count)}
            else:
                h.shard_prepare_send(v_off[self.rank])
                b = h.shard_buffers()
                self._exchange(torch, dev, [(b["positions"], 3 * counts[self.rank][3], "<f4"), (b["normals"], 3 * counts[self.rank][3], "<f4"),
                                            (b["triangle_vertex_ids"], 3 * counts[self.rank][4], "<i4")], None)
        tm.append(time.perf_counter())
        self.last_gpu_ms = gpu_ms
        self.last_fallback = fallback
        # host-side phase times of the last step (ms): local shard + local weld, counts / ranges / boundary keys, key gather + resolve,
        # rest of the shard gather + fix-up (root-weld fallback: gather + weld)
        self.last_phases = [round((b - a) * 1e3, 3) for a, b in zip(tm[:-1], tm[1:])]
        return out

    def step(self):
        return self._step_single() if self.world == 1 else self._step_sharded()

    # -- end-to-end arm: host scene in, pinned host mesh out, every step -----------------------------------
    def e2e(self, scene, steps, warmup, barrier):
        """Every step: scene table host -> device (sdm_set_scene), remesh, mesh device -> pinned host memory.  The download is
        issued asynchronously (sdm_mesh_download_async) into one of two pinned host buffer sets and overlaps the next step's
        compute; all downloads have landed before the clock stops."""
        import torch

        h = self.h
        scene = np.ascontiguousarray(scene)
        h2d = int(scene.nbytes)
        d2h = 0
        host = [None, None]
        barrier()
        t0 = time.perf_counter()
        for i in range(warmup + steps):
            if i == warmup:
                if self.rank == 0:
                    h.download_wait()
                barrier()
                t0 = time.perf_counter()
            h.set_scene(scene)                 # host -> device: the scene table (the path's only input)
            self.step()
            if self.rank == 0:
                m = self.mesh
                need_v, need_t = int(m.vertex_count), int(m.triangle_count)
                hb = host[i & 1]
                if hb is None or hb[0].shape[0] < need_v or hb[2].shape[0] < need_t:
                    h.download_wait()
                    hb = (torch.empty((max(need_v, 1), 3), dtype=torch.float32).pin_memory(),
                          torch.empty((max(need_v, 1), 3), dtype=torch.float32).pin_memory(),
                          torch.empty((max(need_t, 1), 3), dtype=torch.int32).pin_memory())
                    host[i & 1] = hb
                h.download_into_async(m, hb[0].data_ptr(), hb[1].data_ptr(), hb[2].data_ptr())   # device -> host: the mesh
                d2h = need_v * 24 + need_t * 12
        if self.rank == 0:
            h.download_wait()
        barrier()
        return {"elapsed": time.perf_counter() - t0, "h2d": h2d, "d2h": d2h}
