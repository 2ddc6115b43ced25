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
            self._halt.wait(0.2)

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
# CPU baseline: bounded sample of the same workload through the host-compiled reference / the oracle port
# ---------------------------------------------------------------------------------------------------------------
def cpu_sampled_remesh(scene_name, scene, bb, init, levels, res, budget_s=15.0, seed=0):
    """Sampled descent, entirely on the host: at every level a random sample of the surviving voxels is refined with
    the reference's refine kernel (timed), which also gives the survival ratio; at the finest level a sample is meshed
    with the reference's mesh kernel and welded (timed).  Per-level voxel counts and the whole-remesh time are
    extrapolated linearly from the samples.  Uses oracle/_ref/libref_host.so (the reference's own kernels, host-compiled)
    for the sd_obj scene, the oracle port otherwise; all host threads."""
    from oracle import oracle as orc

    use_ref = scene_name == "sd_obj" and orc.RefHost.available()
    o = orc.Oracle(scene)
    ref = orc.RefHost() if use_ref else None
    threads = o.threads()
    rng = np.random.default_rng(seed)

    def refine_fn(v, s):
        return ref.refine_raw(v, s) if use_ref else o.refine_raw(v, s)

    def mesh_fn(v, s):
        return ref.mesh_raw(v, s) if use_ref else o.mesh_raw(v, s)[0]

    def pick(lst, k):
        if lst.shape[0] <= k:
            return lst
        return np.ascontiguousarray(lst[np.sort(rng.choice(lst.shape[0], k, replace=False))])

    vox, vs = o.create_voxel_field(bb, init)
    est_count = float(vox.shape[0])
    total = 0.0
    parts = []
    per_level_budget = budget_s * 0.35 / max(levels, 1)
    for l in range(levels):
        probe = pick(vox, 512)
        t = time.perf_counter(); refine_fn(probe, vs); rate = probe.shape[0] / max(time.perf_counter() - t, 1e-6)
        smp = pick(vox, int(max(4096, rate * per_level_budget)))
        t = time.perf_counter(); raw = refine_fn(smp, vs); dt = time.perf_counter() - t
        keep = np.isfinite(raw).all(axis=1)
        children = raw[keep]
        total += dt * est_count / smp.shape[0]
        parts.append(f"refine L{l}: {smp.shape[0]} of ~{est_count:.0f} voxels in {dt:.2f}s")
        est_count *= children.shape[0] / smp.shape[0]
        vox, vs = np.ascontiguousarray(children), (vs / np.float32(2)).astype(np.float32)
        if vox.shape[0] == 0:
            break
    tri_total = 0.0
    if vox.shape[0]:
        probe = pick(vox, 256)
        t = time.perf_counter(); mesh_fn(probe, vs); rate = probe.shape[0] / max(time.perf_counter() - t, 1e-6)
        smp = pick(vox, int(max(512, rate * budget_s * 0.6)))
        t = time.perf_counter(); tris = mesh_fn(smp, vs); t_mesh = time.perf_counter() - t
        t = time.perf_counter(); pos, nrm, idx = o.weld(tris); t_weld = time.perf_counter() - t
        total += (t_mesh + t_weld) * est_count / smp.shape[0]
        tri_total = idx.shape[0] * est_count / smp.shape[0]
        parts.append(f"mesh kernel + weld: {smp.shape[0]} of ~{est_count:.0f} finest-level voxels in {t_mesh + t_weld:.2f}s")
    return {
        "value": float(res) ** 3 / total,
        "unit": "samples/s",
        "cores": threads,
        "kind": "reference" if use_ref else "port",
        "sample": "sampled descent on the host; " + "; ".join(parts) + "; counts and times extrapolated linearly",
        "extrapolated_s_per_remesh": total,
        "estimated_finest_voxels": est_count,
        "triangles_per_s": tri_total / total,
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
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    scene_name, bb, init, levels, desc = WORKLOADS[args.workload]
    res = init << levels
    scene = make_scene(scene_name)
    metric = f"effective SDF samples/s per full remesh @{res}^3"
    config = {"workload": args.workload, "description": desc, "scene": scene_name, "primitives": int(scene.shape[0]),
              "bb_size": bb, "init_factor": init, "levels": levels, "resolution": res}

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

    runner = parallel.ShardedRemesher(h, bb, init, levels, rank, world, dist)

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
        stage_of = {"k_refine": "refine", "k_cases+k_tri_offsets": "classify", "k_project": "project", "k_vertex_normals": "normals", "k_orient": "orient"}
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
            "config": dict(config, l2="every step clears >0.5 GB of hash tables and rewrites all intermediates (working set > 126 MB L2); no separate flush",
                           parallelism=f"x-slab shards of the level-{runner.split_level} active list over {world} GPU(s), mesh shards gathered to rank 0 (NCCL)" if world > 1 else "1 GPU"),
            "triangles_per_s": tri_count * args.steps / elapsed, "triangles": tri_count, "vertices": vert_count,
            "gpu_ms_per_step": gpu_ms / args.steps, "sdf_evals_per_step": st["sdf_evals"], "sdf_evals_per_s": st["sdf_evals"] * args.steps / elapsed,
            "level_counts": st["level_counts"][: levels + 1], "prim_point_evals_per_step": st["prim_evals"], "gpu_launches": launches, "clocks": clocks,
            "e2e": {"value": float(res) ** 3 * args.steps / e2e["elapsed"], "unit": "samples/s", "h2d_bytes_per_step": e2e["h2d"],
                    "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e["elapsed"] * 1e3 / args.steps},
            "kernel_ms": {k: round(v, 5) for k, v in kavg.items()},
            "roofline": roofline, "roofline_hbm": roofline_hbm,
        }
        if world > 1:
            line["rank0_phase_ms"] = dict(zip(("local_shard_and_weld", "counts_ranges_boundary_keys", "resolve", "gather"), getattr(runner, "last_phases", [])))
            line["root_weld_fallback"] = bool(getattr(runner, "last_fallback", False))
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sampled_remesh(scene_name, scene, bb, init, levels, res)
        print(json.dumps(line), flush=True)
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
    wall, gpu, tris = [], [], []
    for f in range(frames):
        t0 = time.perf_counter()
        h.set_scene(tables[f])
        m = h.remesh(bb, init, levels, download=False)
        wall.append((time.perf_counter() - t0) * 1e3)
        gpu.append(h.stats()["last_gpu_ms"])
        tris.append(int(m.triangle_count))
    w = np.sort(np.asarray(wall)); g = np.sort(np.asarray(gpu))
    pct = lambda a, p: float(a[min(len(a) - 1, int(round(p / 100.0 * (len(a) - 1))))])
    line = {"metric": f"per-frame remesh latency @{res}^3 (animated scene)", "value": pct(w, 50), "unit": "ms", "n_gpus": 1, "steps": frames, "warmup": 3,
            "ms_per_step": float(np.mean(w)), "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic procedural scene (analytic SDF); no dataset", "config": config,
            "latency_ms": {"p50": pct(w, 50), "p90": pct(w, 90), "p99": pct(w, 99), "max": float(w[-1]), "mean": float(np.mean(w))},
            "gpu_latency_ms": {"p50": pct(g, 50), "p99": pct(g, 99), "max": float(g[-1])},
            "triangles_per_frame": {"min": int(min(tris)), "max": int(max(tris))}, "gpu_launches": h.stats()["kernel_launches"]}
    print(json.dumps(line), flush=True)
    h.close()


def run_reference(args, scene_name, scene, bb, init, levels, res, metric, config):
    """--impl reference: the reference's own CPU code path for this workload (oracle/_ref for sd_obj, else the oracle
    port), all host threads; each step is one bounded sampled remesh (cpu_sampled_remesh) extrapolated to a full one."""
    from oracle import oracle as orc

    orc.build()
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    results = [cpu_sampled_remesh(scene_name, scene, bb, init, levels, res, budget_s=12.0, seed=i) for i in range(warm + steps)]
    r = results[-1]
    value = float(np.mean([x["value"] for x in results[-steps:]]))
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": float(res) ** 3 / value * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic procedural scene (analytic SDF); no dataset", "config": config,
        "triangles_per_s": r["triangles_per_s"],
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
