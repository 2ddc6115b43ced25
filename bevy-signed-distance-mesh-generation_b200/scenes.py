"""Scene tables (SdmPrimitive arrays, include/sdfmesh.h) for the configurations of BASELINE.json.

The reference hard-codes a single scene, ``sd_obj`` (cuda/modules/common.cu:222-226); every other
configuration is assembled from the reference's primitive forms in cuda/includes/signed_distance.cu
(SURVEY.md section 8d, C1-C5).  A scene is a numpy structured array so that the very same bytes are handed
to the CUDA library and, in the tests, to the CPU oracle.
"""
from __future__ import annotations

import numpy as np

PRIM_DTYPE = np.dtype(
    [("kind", "<u4"), ("fold", "<u4"), ("k", "<f4"), ("radius", "<f4"), ("a", "<f4", (3,)), ("b", "<f4", (3,))]
)
assert PRIM_DTYPE.itemsize == 40

SPHERE, BOX, CAPSULE, BOX_SKELETON, MANDELBULB = 0, 1, 2, 3, 4
FOLD_MIN, FOLD_SMOOTH_MIN = 0, 1


def _prim(kind, fold=FOLD_MIN, k=0.0, radius=0.0, a=(0, 0, 0), b=(0, 0, 0)):
    p = np.zeros((), dtype=PRIM_DTYPE)
    p["kind"], p["fold"], p["k"], p["radius"] = kind, fold, np.float32(k), np.float32(radius)
    p["a"], p["b"] = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return p


def sd_obj() -> np.ndarray:
    """common.cu:222-226: smooth_min(sd_box_skeleton(p, 0, (3,1,.5), .1), length(p) - 1, .5)."""
    return np.stack(
        [
            _prim(BOX_SKELETON, FOLD_MIN, radius=0.1, a=(0, 0, 0), b=(3.0, 1.0, 0.5)),
            _prim(SPHERE, FOLD_SMOOTH_MIN, k=0.5, radius=1.0, a=(0, 0, 0)),
        ]
    )


def render_scene() -> np.ndarray:
    """sd_scene of the ray-march viewer (compute_render.cu:3-19): min(sd_obj, wire box of the meshing domain with lw 0.05)."""
    return np.concatenate([sd_obj(), np.stack([_prim(BOX_SKELETON, FOLD_MIN, radius=0.05, a=(0, 0, 0), b=(5.0, 5.0, 5.0))])])


def sphere_box() -> np.ndarray:
    """C1: min(length(p-(0.6,0,0)) - 1, sd_box(p, (-0.6,0,0), (1.5,1.5,1.5))) (signed_distance.cu:82-91 forms)."""
    return np.stack(
        [
            _prim(SPHERE, FOLD_MIN, radius=1.0, a=(0.6, 0, 0)),
            _prim(BOX, FOLD_MIN, a=(-0.6, 0, 0), b=(1.5, 1.5, 1.5)),
        ]
    )


def mandelbulb() -> np.ndarray:
    """C4: sd_unit_mandelbulb (signed_distance.cu:55-57): sd_mandelbulb(p / 0.4, 0) * 0.4."""
    return np.stack([_prim(MANDELBULB, FOLD_MIN, radius=0.4)])


def many_primitives(n: int = 1024, seed: int = 1234, k: float = 0.1, t: float | None = None, anim_seed: int = 4321) -> np.ndarray:
    """C3 (and C5 when ``t`` is given): n primitives folded with smooth_min(acc, d_i, k) in index order.

    type = i mod 3: sphere r in U[.05,.15]; capsule |b1-b0| in U[.1,.4], lw in U[.03,.08]; box size in U[.1,.3]^3;
    centres U[-2,2]^3.  With ``t`` the centres move as c_i + d_i*sin(2*pi*t/T_i), the cyclic
    motion of the reference's example scene (src/example_scene.rs:131-144), d_i in U[0,.3]^3, T_i in U[2,8] s.
    """
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-2.0, 2.0, size=(n, 3))
    radius = rng.uniform(0.05, 0.15, size=n)
    seg_len = rng.uniform(0.1, 0.4, size=n)
    seg_dir = rng.normal(size=(n, 3))
    seg_dir /= np.linalg.norm(seg_dir, axis=1, keepdims=True)
    lw = rng.uniform(0.03, 0.08, size=n)
    box = rng.uniform(0.1, 0.3, size=(n, 3))
    if t is not None:
        arng = np.random.default_rng(anim_seed)
        amp = arng.uniform(0.0, 0.3, size=(n, 3))
        period = arng.uniform(2.0, 8.0, size=n)
        centres = centres + amp * np.sin(2.0 * np.pi * t / period)[:, None]
    out = np.zeros(n, dtype=PRIM_DTYPE)
    i = np.arange(n)
    sph, cap, box_ = (i % 3 == 0), (i % 3 == 1), (i % 3 == 2)
    half = 0.5 * seg_len[:, None] * seg_dir
    out["fold"] = FOLD_SMOOTH_MIN
    out["k"] = np.float32(k)
    out["kind"] = np.where(sph, SPHERE, np.where(cap, CAPSULE, BOX))
    out["radius"] = np.where(sph, radius, np.where(cap, lw, 0.0)).astype(np.float32)
    out["a"] = np.where(cap[:, None], centres - half, centres).astype(np.float32)
    out["b"] = np.where(cap[:, None], centres + half, np.where(box_[:, None], box, 0.0)).astype(np.float32)
    return out


SCENES = {
    "sd_obj": sd_obj,
    "sphere_box": sphere_box,
    "mandelbulb": mandelbulb,
    "many1024": many_primitives,
}
