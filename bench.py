#!/usr/bin/env python3
"""bench.py - whole-remesh throughput of the mesh-generation hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A *step* is one full remesh of the workload's scene: level-0 field -> `levels` refinements (SDF lattice
classification + stable compaction) -> marching-cubes classification -> vertex projection, normals, orientation ->
reference-order weld, leaving positions / normals / indices in HBM.  Metric (BASELINE.json): effective SDF
samples/s = R^3 / t_remesh (dense-grid equivalent of the sparse hierarchical evaluation), plus triangles/s and
ms per remesh in the same JSON line.

`value`   : inputs (the compiled scene table) already resident in HBM, result left in HBM.
`e2e`     : the same remesh through the host-buffer C ABI: scene table uploaded from host memory every step and the
            mesh downloaded into pinned host buffers every step (H2D / D2H inside the timed region).
`roofline`: dominant kernel (k_project, FP32-bound; no tensor-core or HBM roofline applies to it) timed live with
            CUDA events on the library's own stream; `roofline_hbm` gives the HBM-bound emit kernel.
`cpu_baseline`: the reference's own kernels host-compiled (oracle/_ref, kind "reference"; sd_obj scenes) or the
            oracle port (other scenes) on a bounded sample of the same workload, extrapolated to a full remesh.

The oracle (oracle/) is only touched by the cpu_baseline leg and by `--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (scene factory name, bb_size, init_factor, levels, description)
    "sd_obj_1024": ("sd_obj", 5.0, 32, 5, "reference scene sd_obj (common.cu:222-226), INIT 32 x 5 levels = 1024^3"),
    "c2_sd_obj_512": ("sd_obj", 5.0, 64, 3, "BASELINE configs[1]: sd_obj, INIT 64 x 3 levels = 512^3"),
    "c3_many1024_1024": ("many1024", 5.0, 64, 4, "BASELINE configs[2]: 1024-primitive smooth-union scene, INIT 64 x 4 levels = 1024^3"),
    "c4_mandelbulb_2048": ("mandelbulb", 5.0, 128, 4, "BASELINE configs[3]: Mandelbulb, INIT 128 x 4 levels = 2048^3"),
    "c1_sphere_box_128": ("sphere_box", 5.0, 32, 2, "BASELINE configs[0]: sphere U box, INIT 32 x 2 levels = 128^3"),
    "c5_animated_1024": ("many1024", 5.0, 64, 4, "BASELINE configs[4]: the 1024-primitive scene with moving centres, remeshed every frame at 1024^3 (p50/p99 latency)"),
}
DEFAULT_WORKLOAD = "c3_many1024_1024"

# Algorithmic FP32 operations per primitive evaluation (add/sub/mul/min/max/compare-select; sqrt and div counted as 1):
# DESIGN.md "Algorithmic work".  Used only to convert evaluations/s into the roofline's TFLOP/s.
OPS = {"capsule": 24, "sphere": 10, "box": 25, "smooth_min": 10, "min": 1, "mandelbulb": 25 * 60}


def scene_ops_per_eval(scene: np.ndarray) -> int:
    from bsdmg_b200 import scenes as S

    total = 0
    for p in scene:
        kind = int(p["kind"])
        fold = OPS["smooth_min"] if int(p["fold"]) == S.FOLD_SMOOTH_MIN else OPS["min"]
        if kind == S.SPHERE:
            total += OPS["sphere"] + fold
        elif kind == S.BOX:
            total += OPS["box"] + fold
        elif kind == S.CAPSULE:
            total += OPS["capsule"] + fold
        elif kind == S.BOX_SKELETON:
            total += 12 * (OPS["capsule"] + OPS["min"])
        elif kind == S.MANDELBULB:
            total += OPS["mandelbulb"] + fold
    return total


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append((float(out[0]), float(out[1])))
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.02)   # an nvidia-smi query takes ~40 ms itself; the default timed region is ~0.2 s

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons), "samples": len(sm)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def make_scene(name):
    from bsdmg_b200 import scenes

    return scenes.SCENES[name]()


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own kernels, host-compiled, on a bounded sub-volume of the same workload
# ---------------------------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Threads the CPU arm uses: every core this process may run on (torchrun exports OMP_NUM_THREADS=1: ignored on purpose)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class CpuPath:
    """The path on the host: refine kernel + stable retain per level, mesh kernel, weld - through oracle/_ref/libref_host.so
    (kind "reference": the reference's kernels compiled unmodified for sd_obj, their functor templates over the reference's
    own primitives for the other scenes, oracle/ref_functor.inc) or, where that library is absent, the oracle port."""

    def __init__(self, scene_name, scene):
        from oracle import oracle as orc

        self.orc = orc
        self.scene = scene
        self.port = orc.Oracle(scene)
        self.threads = host_threads()
        orc.Oracle.set_threads(self.threads)
        self.ref = orc.RefHost() if orc.RefHost.available() else None
        self.kind = "reference" if self.ref is not None else "port"
        if self.ref is not None:
            self.ref.set_threads(self.threads)
        self.mode = 0 if scene_name == "sd_obj" else (2 if scene_name == "mandelbulb" else 3)

    def refine(self, vox, vs):
        if self.ref is None:
            raw = self.port.refine_raw(vox, vs)
        elif self.mode == 0:
            raw = self.ref.refine_raw(vox, vs)
        else:
            raw = self.ref.tpl_refine_raw(self.mode, self.scene, vox, vs)
        keep = np.isfinite(raw).all(axis=1)                      # Vec::retain (src/cuda/mod.rs:192-193), stable
        return np.ascontiguousarray(raw[keep]), (vs / np.float32(2)).astype(np.float32)

    def mesh(self, vox, vs):
        if self.ref is None:
            tris = self.port.mesh_raw(vox, vs)[0]
        elif self.mode == 0:
            tris = self.ref.mesh_raw(vox, vs)
        else:
            tris = self.ref.tpl_mesh_raw(self.mode, self.scene, vox, vs)
        return self.orc.Oracle.weld(tris)                          # src/cuda/mod.rs:263-296 (Rust host step, restated)

    def remesh_cells(self, cells, vs, levels):
        """The complete path on a sub-volume: the given level-0 voxels -> `levels` refinements -> mesh -> weld."""
        vox = cells
        for _ in range(levels):
            if vox.shape[0] == 0:
                break
            vox, vs = self.refine(vox, vs)
        if vox.shape[0] == 0:
            return 0, 0
        pos, _, idx = self.mesh(vox, vs)
        return int(idx.shape[0]), int(vox.shape[0])


def cpu_subvolume_steps(scene_name, scene, bb, init, levels, res, steps, warmup, target_s=5.0):
    """Each step = the COMPLETE path (all refinement levels, mesh kernel, weld) on a uniformly random subset of the level-0
    voxels - a bounded sub-volume of the same workload - sized by a calibration probe so that one step takes ~target_s on
    this host.  Effective samples of a step = (its share of the level-0 voxels) x res^3, the dense-grid equivalent the GPU
    arm's metric uses; nothing is extrapolated inside a step, the step time is what was measured."""
    cpu = CpuPath(scene_name, scene)
    vox0, vs0 = cpu.port.create_voxel_field(bb, init)
    n0 = vox0.shape[0]
    rng = np.random.default_rng(2024)

    def pick(k):
        return np.ascontiguousarray(vox0[np.sort(rng.choice(n0, min(k, n0), replace=False))])

    # calibration: grow the subset until it takes a measurable time, then scale to the target
    k = min(n0, 64)
    while True:
        t = time.perf_counter(); cpu.remesh_cells(pick(k), vs0, levels); dt = time.perf_counter() - t
        if dt >= 0.5 or k >= n0:
            break
        k = min(n0, k * 4)
    k = int(max(8, min(n0, k * target_s / max(dt, 1e-3))))
    times, tris = [], []
    for i in range(warmup + steps):
        cells = pick(k)
        t = time.perf_counter(); nt, _ = cpu.remesh_cells(cells, vs0, levels); dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt); tris.append(nt)
    total = float(sum(times))
    share = k / n0
    eff = share * float(res) ** 3
    return {
        "value": eff * len(times) / total, "unit": "samples/s", "cores": cpu.threads, "kind": cpu.kind,
        "sample": (f"each step = the complete path (level-0 voxels -> {levels} refinements -> mesh kernel -> weld) on {k} of {n0} level-0 voxels "
                   f"(uniform random subset, 1/{n0 / k:.1f} of the domain); effective samples per step = {k}/{n0} x {res}^3; "
                   f"a whole remesh at this rate would take {total / len(times) / share:.0f} s"),
        "ms_per_step": total / len(times) * 1e3, "triangles_per_s": float(sum(tris)) / total,
        "level0_voxels_per_step": k, "level0_voxels": n0, "whole_remesh_s_at_this_rate": total / len(times) / share,
    }


def ncu_traffic(workload, kernels):
    """dram__bytes_read.sum + dram__bytes_write.sum per remesh of the given kernels, from the committed `ncu --set full` capture of
    this workload (profiles/ncu_traffic.json, see its _doc); None if that workload was not captured."""
    try:
        t = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text()).get(workload)
        return float(sum(t[k]["dram_bytes_read"] + t[k]["dram_bytes_write"] for k in kernels)) if t else None
    except (OSError, KeyError, ValueError):
        return None


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU exchange: device-driven over peer memory, or host-driven over NCCL")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    scene_name, bb, init, levels, desc = WORKLOADS[args.workload]
    res = init << levels
    scene = make_scene(scene_name)
    metric = f"effective SDF samples/s per full remesh @{res}^3"
    from bsdmg_b200 import parallel as _par

    config = {"workload": args.workload, "description": desc, "scene": scene_name, "primitives": int(scene.shape[0]),
              "bb_size": bb, "init_factor": init, "levels": levels, "resolution": res,
              "l2": "every step clears >0.3 GB of tables / bitmaps and rewrites all intermediates (working set > 126 MB L2); no separate flush",
              "parallelism": (f"x-slab shards of the level-{_par.choose_split_level(init, levels, max(world, args.gpus), _par.scene_is_culled(scene))} active list over {max(world, args.gpus)} GPU(s), "
                              "welded where they are made, interface keys resolved on rank 0, rows pushed into rank 0's HBM over NVLink through peer-mapped memory "
                              "(device-side flags; --exchange nccl: the host-driven NCCL exchange)") if max(world, args.gpus) > 1 else "1 GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference(args, scene_name, scene, bb, init, levels, res, metric, config)
        return
    if args.workload == "c5_animated_1024":
        run_animated(args, bb, init, levels, res, config)
        return

    import torch
    import bsdmg_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "16")   # the gather is point-to-point: 421 -> 492 GB/s into rank 0 (tools/p2p_probe.py)
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    h = bsdmg_b200.CudaHandler(local_rank, scene)

    from bsdmg_b200 import parallel

    exchange = "single"
    if world > 1 and args.exchange == "peer":
        try:   # device-driven exchange over peer-mapped memory (CUDA IPC); the host-driven NCCL exchange remains as the fallback
            runner = parallel.PeerRemesher(h, bb, init, levels, rank, world, dist, culled=parallel.scene_is_culled(scene))
            exchange = "peer (device-side flags, stores over NVLink into rank 0 / per-rank PCIe for e2e)"
        except Exception as exc:   # noqa: BLE001 - any set-up failure (IPC not permitted, ...) must not lose the measurement
            print(f"[rank {rank}] peer exchange unavailable ({exc}); using the NCCL exchange", file=sys.stderr, flush=True)
            runner = parallel.ShardedRemesher(h, bb, init, levels, rank, world, dist, culled=parallel.scene_is_culled(scene))
            exchange = "nccl (host-driven)"
    else:
        runner = parallel.ShardedRemesher(h, bb, init, levels, rank, world, dist, culled=parallel.scene_is_culled(scene))
        if world > 1:
            exchange = "nccl (host-driven)"

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ---------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        out = runner.step()
    launches0 = h.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    gpu_ms = 0.0
    for _ in range(args.steps):
        out = runner.step()
        gpu_ms += runner.last_gpu_ms
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop()
    launches = h.stats()["kernel_launches"] - launches0
    elapsed = t1 - t0
    if dist is not None:
        tt = torch.tensor([elapsed, gpu_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed, gpu_ms = float(tt[0]), float(tt[1])
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])
    ms_per_step = elapsed * 1e3 / args.steps
    tri_count, vert_count = out["triangles"], out["vertices"]

    # ---- per-kernel times (profiling pass, outside the timed region) -------------------------------------------
    h.set_profiling(True)
    ktimes = {}
    reps = 5
    for _ in range(reps):
        runner.step()
        for name, ms in h.kernel_times():
            ktimes.setdefault(name, []).append(ms)
    h.set_profiling(False)
    st = h.stats()
    try:   # DevState word 55: triangles the six-sample orientation test left to the twelve-sample pass (this rank's shard)
        orient_pending = int(h.debug_fetch("state", 76, np.uint32)[55])
    except Exception:
        orient_pending = None
    kavg = {k: (sum(v) / reps) for k, v in ktimes.items()}           # ms per step, summed over the launches of that name
    step_sum = sum(kavg.values())
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    ops_per_eval = scene_ops_per_eval(scene)
    fp32_peak_tflops = 148 * 128 * sm_max_mhz * 1e6 / 1e12         # non-FMA: parity requires -fmad=false
    roofline = roofline_hbm = None
    if rank == 0 and kavg:
        # dominant SDF kernel: achieved = (primitive, point) distance evaluations it actually folded (device counters, after
        # culling) x algorithmic FP32 ops per such evaluation (scene average) / its CUDA-event time
        pe = st["prim_evals"]
        stage_of = {"k_refine": "refine", "k_cases+k_tri_offsets": "classify", "k_project": "project", "k_project_tail": "tail",
                    "k_vertex_normals": "normals", "k_orient": "orient"}
        cand = {k: v for k, v in kavg.items() if k in stage_of}
        top = max(cand, key=cand.get)
        nprims_compiled = sum(12 if int(p["kind"]) == 3 else 1 for p in scene)
        ops_pp = ops_per_eval / max(nprims_compiled, 1)
        work = pe[stage_of[top]]
        ach = work * ops_pp / (kavg[top] * 1e-3) / 1e12
        roofline = {"kernel": top, "bound": "fp32", "achieved": ach, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                    "frac": ach / fp32_peak_tflops, "traffic": ncu_traffic(args.workload, [top]),
                    "peak_source": f"148 SM x 128 FP32 lanes x {sm_max_mhz:.0f} MHz, one op per lane-cycle (no FMA: -fmad=false is part of the parity contract)",
                    "algorithmic_ops_per_prim_point": ops_pp, "prim_point_evals_per_launch": work, "avg_launch_ms": kavg[top],
                    "share_of_step": kavg[top] / step_sum}
        # k_emit_vertices: per vertex first_slot 4 + two bit-rank words 8 + position/normal in 24 and out 24 + output index 4;
        # k_emit_indices: per triangle three vertex ids 12 + three output-index gathers 12 + three indices out 12
        emit_bytes = vert_count * (4 + 8 + 24 + 24 + 4) + tri_count * (12 + 12 + 12)
        if "k_emit_vertices" in kavg and "k_emit_indices" in kavg:
            t_emit = (kavg["k_emit_vertices"] + kavg["k_emit_indices"]) * 1e-3
            roofline_hbm = {"kernel": "k_emit_vertices+k_emit_indices", "bound": "hbm", "achieved": emit_bytes / t_emit / 1e9, "peak": hbm_peak,
                            "unit": "GB/s", "frac": emit_bytes / t_emit / 1e9 / hbm_peak, "traffic": ncu_traffic(args.workload, ["k_emit_vertices", "k_emit_indices"]),
                            "peak_source": peak_src,
                            "algorithmic_bytes": emit_bytes}

    # ---- per-rank view (multi-GPU): shard sizes and device times, to see the balance of the split
    per_rank = None
    if dist is not None:
        mine = {"rank": rank, "finest_voxels": st["level_counts"][levels], "unique_vertices": st["unique_vertices"],
                "gpu_ms": round(float(sum(kavg.values())), 4),
                "kernel_ms": {k: round(v, 4) for k, v in kavg.items() if v > 0.02}}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered

    # ---- checksum of the mesh the timed loop produced (outside the timed region): FNV-1a-64 of the raw bytes, the checksum of
    #      tests/golden; identical at every N because the merged mesh is byte-identical to the single-GPU mesh
    mesh_fnv = None
    if rank == 0 and runner.mesh is not None:
        mm = h._download(runner.mesh)
        lib = bsdmg_b200.load_library()
        fnv = lambda a: "%016x" % int(lib.sdm_hash_bytes(a.ctypes.data, a.nbytes))
        mesh_fnv = {"indices": fnv(mm.indices), "positions": fnv(mm.positions), "normals": fnv(mm.normals)}
        del mm

    # ---- end-to-end arm: host scene in, pinned host mesh out, every step ------------------------------------------
    e2e = runner.e2e(scene, args.steps, max(args.warmup, 3), barrier)
    if dist is not None:
        et = torch.tensor([e2e["elapsed"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        e2e["elapsed"] = float(et[0])

    if rank == 0:
        line = {
            "metric": metric, "value": float(res) ** 3 * args.steps / elapsed, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic procedural scene (analytic SDF); no dataset",
            "config": config,
            "triangles_per_s": tri_count * args.steps / elapsed, "triangles": tri_count, "vertices": vert_count, "mesh_fnv": mesh_fnv,
            "gpu_ms_per_step": gpu_ms / args.steps, "sdf_evals_per_step": st["sdf_evals"], "sdf_evals_per_s": st["sdf_evals"] * args.steps / elapsed,
            "level_counts": st["level_counts"][: levels + 1], "prim_point_evals_per_step": st["prim_evals"], "gpu_launches": launches, "clocks": clocks,
            "newton": {"iterations": st["newton_iterations"], "stragglers": st["stragglers"], "escaped_vertices": st["escaped_vertices"],
                       "list_fallback_tiles": st["list_fallback_tiles"]},
            "orient_pending_triangles": orient_pending,
            "e2e": {"value": float(res) ** 3 * args.steps / e2e["elapsed"], "unit": "samples/s", "h2d_bytes_per_step": e2e["h2d"],
                    "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e["elapsed"] * 1e3 / args.steps},
            "kernel_ms": {k: round(v, 5) for k, v in kavg.items()},
            "roofline": roofline, "roofline_hbm": roofline_hbm,
        }
        if world > 1:
            line["exchange"] = exchange
            line["per_rank"] = per_rank
            line["root_weld_fallback"] = bool(getattr(runner, "last_fallback", False))
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_subvolume_steps(scene_name, scene, bb, init, levels, res, steps=3, warmup=0, target_s=5.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "triangles_per_s")}
        print(json.dumps(line), flush=True)
    if hasattr(runner, "close"):
        runner.close()
    h.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_animated(args, bb, init, levels, res, config):
    """BASELINE configs[4]: every frame gets a new scene table (centres move as in src/example_scene.rs:131-144), uploaded
    from the host, its culling masks rebuilt, and a full remesh; per-frame latency is the wall time of set_scene + remesh
    (result left in HBM).  Frames = --steps (default 600), t = frame / 60."""
    import bsdmg_b200
    from bsdmg_b200 import scenes

    frames = args.steps if args.steps != 20 else 600
    tables = [scenes.many_primitives(1024, t=f / 60.0) for f in range(frames)]
    h = bsdmg_b200.CudaHandler(0, tables[0])
    for f in range(3):
        h.set_scene(tables[f]); h.remesh(bb, init, levels, download=False)
    wall, gpu, tris, strag, iters = [], [], [], [], []
    for f in range(frames):
        t0 = time.perf_counter()
        h.set_scene(tables[f])
        m = h.remesh(bb, init, levels, download=False)
        wall.append((time.perf_counter() - t0) * 1e3)
        st = h.stats()
        gpu.append(st["last_gpu_ms"]); strag.append(st["stragglers"]); iters.append(st["newton_iterations"])
        tris.append(int(m.triangle_count))
    # frames far above the median, explained: the same frame again with per-kernel timing (outside the latency statistics)
    med = float(np.median(wall))
    outliers = []
    h.set_profiling(True)
    for f in [int(i) for i in np.argsort(wall)[::-1][:3] if wall[int(i)] > 2.0 * med]:
        h.set_scene(tables[f]); h.remesh(bb, init, levels, download=False)
        kt = {}
        for name, ms in h.kernel_times():
            kt[name] = kt.get(name, 0.0) + ms
        top = sorted(kt.items(), key=lambda kv: -kv[1])[:3]
        outliers.append({"frame": f, "t": f / 60.0, "wall_ms": round(wall[f], 3), "again_gpu_ms": round(h.stats()["last_gpu_ms"], 3), "stragglers": strag[f],
                         "newton_iterations": iters[f], "top_kernels_ms": {k: round(v, 3) for k, v in top}})
    h.set_profiling(False)
    w = np.sort(np.asarray(wall)); g = np.sort(np.asarray(gpu))
    pct = lambda a, p: float(a[min(len(a) - 1, int(round(p / 100.0 * (len(a) - 1))))])
    line = {"metric": f"per-frame remesh latency @{res}^3 (animated scene)", "value": pct(w, 50), "unit": "ms", "n_gpus": 1, "steps": frames, "warmup": 3,
            "ms_per_step": float(np.mean(w)), "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic procedural scene (analytic SDF); no dataset", "config": config,
            "latency_ms": {"p50": pct(w, 50), "p90": pct(w, 90), "p99": pct(w, 99), "max": float(w[-1]), "mean": float(np.mean(w))},
            "gpu_latency_ms": {"p50": pct(g, 50), "p99": pct(g, 99), "max": float(g[-1])},
            "triangles_per_frame": {"min": int(min(tris)), "max": int(max(tris))}, "gpu_launches": h.stats()["kernel_launches"],
            "newton": {"stragglers_per_frame": {"p50": int(np.median(strag)), "max": int(max(strag))},
                       "iterations_per_frame": {"p50": int(np.median(iters)), "max": int(max(iters))}},
            "outliers": outliers}
    print(json.dumps(line), flush=True)
    h.close()


def run_reference(args, scene_name, scene, bb, init, levels, res, metric, config):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref/libref_host.so: its kernels compiled
    for the host, all host threads) on this arm's workload, metric and unit.  Every one of the --steps K timed steps (after
    --warmup W) is a bounded sample of the workload: the complete path on a random sub-volume sized to ~5 s, so that
    `ms_per_step` x steps is the time this run really takes."""
    from oracle import oracle as orc

    orc.build()
    r = cpu_subvolume_steps(scene_name, scene, bb, init, levels, res, steps=max(1, args.steps), warmup=max(0, args.warmup), target_s=5.0)
    line = {
        "impl": "reference", "metric": metric, "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": max(1, args.steps),
        "warmup": max(0, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic procedural scene (analytic SDF); no dataset", "config": config,
        "triangles_per_s": r["triangles_per_s"],
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "whole_remesh_s_at_this_rate": r["whole_remesh_s_at_this_rate"], "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
