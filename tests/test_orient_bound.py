"""CPU property test of the claim behind k_orient<true> (csrc/sdm_kernels.cuh, DESIGN.md "Orientation from six samples"), on the
oracle's evaluation of real scenes: with d_axis = 8 A_axis - B_axis (signed_distance.cu:186-199; A = f(+e) - f(-e), B = f(+2e) -
f(-2e)), the outer differences obey |B_axis| <= 4e (1-Lipschitz scene, plus rounding), and whenever the kernel's test
7.9 |tn . A| > |tn|_1 * Bmax holds, sign(tn . d) = sign(tn . A) - for the reference's own float evaluation of d."""
import numpy as np
import pytest

from bsdmg_b200 import scenes

EPS = np.float32(0.001)   # signed_distance.cu:179


def _samples(orc, pts):
    """f at the 12 stencil points of empirical_normal, built as the reference builds them (p + vec3(off, 0, 0) ...): [n, 3 axes, 4]"""
    offs = np.array([2.0 * EPS, EPS, -EPS, -2.0 * EPS], np.float32)
    out = np.empty((pts.shape[0], 3, 4), np.float32)
    for a in range(3):
        for s in range(4):
            q = pts.copy()
            q[:, a] = q[:, a] + offs[s]
            out[:, a, s] = orc.sdf(q)
    return out


def _reach(table):
    """the host's bound (compile_scene): max over the table of |centre|_1 + extent, + kmax + 1; generous for this test"""
    return 16.0


@pytest.mark.parametrize("scene_name,n", [("many256", 6000), ("sd_obj", 20000)])
def test_six_samples_decide_the_orientation_sign(oracle_mod, scene_name, n):
    table = scenes.many_primitives(256) if scene_name == "many256" else scenes.sd_obj()
    orc = oracle_mod.Oracle(table)
    rng = np.random.default_rng(7)
    # points near the surface (where triangle centroids are): random points pulled onto it, then jittered by up to a voxel
    p0 = rng.uniform(-2.2, 2.2, size=(n, 3)).astype(np.float32)
    proj, _ = orc.project(p0)
    ok = np.isfinite(proj).all(axis=1)
    pts = (proj[ok] + rng.uniform(-5e-3, 5e-3, size=(int(ok.sum()), 3))).astype(np.float32)
    pts = np.concatenate([pts, p0[: n // 4]])   # and anywhere in the domain
    f = _samples(orc, pts)
    f0, f1, f2, f3 = f[:, :, 0], f[:, :, 1], f[:, :, 2], f[:, :, 3]
    # the reference's left-to-right float sum (signed_distance.cu:190-198)
    d = (((-f0) + np.float32(8.0) * f1) - np.float32(8.0) * f2) + f3
    A = (f1 - f2).astype(np.float64)
    B = (f0 - f3).astype(np.float64)
    fin = np.isfinite(f).all(axis=(1, 2))
    assert fin.mean() > 0.99
    # 1-Lipschitz: the outer samples are 4e apart
    assert np.abs(B[fin]).max() <= 4.0 * float(EPS) * 1.003 + 2e-5
    # face normals: random unit vectors, axis-aligned ones, and the scene's own normal rotated away by up to 90 degrees
    tn = rng.normal(size=pts.shape).astype(np.float64)
    tn /= np.linalg.norm(tn, axis=1, keepdims=True)
    tn[::7] = np.eye(3)[rng.integers(0, 3, size=tn[::7].shape[0])] * rng.choice([-1.0, 1.0], size=(tn[::7].shape[0], 1))
    tn = tn.astype(np.float32).astype(np.float64)
    dotA = (tn * A).sum(axis=1)
    l1 = np.abs(tn).sum(axis=1)
    fmax = np.abs(f[:, :, 1:3]).max(axis=(1, 2)).astype(np.float64)
    L = table.shape[0] * 12 if scene_name == "sd_obj" else table.shape[0]   # folded primitives (un-culled: the whole compiled table)
    unit = 2e-7 * (_reach(table) + np.abs(pts.astype(np.float64)).sum(axis=1))
    bmax = 4.0 * float(EPS) * 1.003 + 2.0 * (min(L, 16) + 6) * unit + 4e-6 * (fmax + 1e-3)   # the kernel's L is the culled list (<= 16 here)
    decided = fin & (7.9 * np.abs(dotA) > l1 * bmax)
    assert decided.mean() > 0.5, "the test should decide most cases"
    dot_d = (tn * d.astype(np.float64)).sum(axis=1)
    assert np.all(np.sign(dot_d[decided]) == np.sign(dotA[decided])), "a decided sign differs from the twelve-sample statement"
    # ... with the margin the kernel's comment claims: |cos(tn, n)| > 1e-3
    cosv = np.abs(dot_d[decided]) / (np.linalg.norm(d[decided].astype(np.float64), axis=1) * np.linalg.norm(tn[decided], axis=1))
    assert cosv.min() > 1e-3
