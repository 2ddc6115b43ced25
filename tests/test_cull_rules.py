"""CPU property tests of the culling rules the CUDA kernels use on large scenes (csrc/sdm_device.cuh, tile_refine_lanes /
tile_refine_halves; DESIGN.md "Culling for large scenes"), restated in numpy and checked on the ORACLE's evaluation:

* exactness: for a ball (c, rho), the primitives the rule keeps (drop test + reset rule, fold order = index order) fold to the SAME
  BITS as the whole table at every point of the ball - the oracle evaluates the whole table and the kept sub-table;
* the candidate-parallel form of the rule (32 candidates at a time, the running minimum as a prefix minimum: the Newton tail
  kernel) decides exactly what the serial form decides, NaN and infinite distances included.
"""
import numpy as np
import pytest

from bsdmg_b200 import scenes

F = np.float32
INF = F(np.inf)


def need_list_serial(d, kk, r, kmax):
    """tile_refine_lanes for one lane: keep unless d - k >= U + A; reset where U - B >= d + k; U = running min of d + r."""
    A, B = F(r + F(1e-4)), F(F(3.0) * r + kmax + F(1e-4))
    U, first, keep = INF, 0, np.zeros(d.shape[0], bool)
    for q in range(d.shape[0]):
        keep[q] = not (F(d[q] - kk[q]) >= F(U + A))     # NaN: keep
        if F(U - B) >= F(d[q] + kk[q]):
            first = q
        U = np.fmin(U, F(d[q] + r))                       # fminf: a NaN distance is skipped
    keep[:first] = False
    return keep, first


def need_list_chunked(d, kk, r, kmax, width=32):
    """tile_refine_halves: `width` candidates per step, exclusive prefix minimum inside the step, carried minimum across steps."""
    A, B = F(r + F(1e-4)), F(F(3.0) * r + kmax + F(1e-4))
    n = d.shape[0]
    Uc, first, keep = INF, 0, np.zeros(n, bool)
    for w0 in range(0, n, width):
        dd = np.full(width, INF, F); k2 = np.zeros(width, F); valid = np.zeros(width, bool)
        m = min(width, n - w0)
        dd[:m], k2[:m], valid[:m] = d[w0:w0 + m], kk[w0:w0 + m], True
        incl = (dd + r).astype(F)
        o = 1
        while o < width:   # Hillis-Steele with fminf, as the shuffles do it
            t = np.concatenate([np.full(o, INF, F), incl[:-o]])
            incl = np.where(np.arange(width) >= o, np.fmin(incl, t), incl)
            o <<= 1
        excl = np.concatenate([[INF], incl[:-1]]).astype(F)
        U = np.fmin(Uc, excl)
        kp = valid & ~((dd - k2).astype(F) >= (U + A).astype(F))
        rs = valid & ((U - B).astype(F) >= (dd + k2).astype(F))
        keep[w0:w0 + m] = kp[:m]
        if rs.any():
            first = w0 + int(np.nonzero(rs)[0].max())
        Uc = np.fmin(Uc, incl[-1])
    keep[:first] = False
    return keep, first


def test_candidate_parallel_rule_equals_serial_rule():
    rng = np.random.default_rng(11)
    for trial in range(300):
        n = int(rng.integers(1, 129))
        d = rng.uniform(-0.2, 3.0, size=n).astype(F)
        if trial % 3 == 0:
            d[rng.integers(0, n, size=max(1, n // 8))] = np.nan
        if trial % 5 == 0:
            d[rng.integers(0, n)] = np.inf
        if trial % 7 == 0:
            d[: rng.integers(1, n + 1)] = np.nan   # leading NaNs: the running minimum starts undefined
        kk = np.where(rng.random(n) < 0.8, F(0.1), F(0.0)).astype(F)
        r = F(rng.choice([0.0022, 0.01, 0.08]))
        with np.errstate(invalid="ignore"):
            a, fa = need_list_serial(d, kk, r, F(0.1))
            for width in (16, 32):
                b, fb = need_list_chunked(d, kk, r, F(0.1), width)
                assert fa == fb and np.array_equal(a, b), (trial, width)


@pytest.mark.parametrize("nprims,rho", [(256, 0.0022), (256, 0.02), (256, 0.1), (1024, 0.0022), (1024, 0.05)])
def test_kept_primitives_fold_to_the_same_bits(oracle_mod, nprims, rho):
    table = scenes.many_primitives(nprims)
    orc = oracle_mod.Oracle(table)
    rng = np.random.default_rng(5)
    n_balls, per_ball = 60, 24
    p0 = rng.uniform(-2.2, 2.2, size=(n_balls, 3)).astype(F)
    proj, _ = orc.project(p0)                                   # centres near the surface: where the kernels evaluate
    c = np.where(np.isfinite(proj), proj, p0).astype(F)
    c[::4] = p0[::4]                                            # ... and anywhere
    # d_i(c) for every primitive alone: a one-primitive table folds FLT_MAX with d_i and returns d_i for both folds
    dist = np.stack([oracle_mod.Oracle(table[i:i + 1]).sdf(c) for i in range(table.shape[0])], axis=1)
    kk = np.where(table["fold"] == scenes.FOLD_SMOOTH_MIN, table["k"], F(0.0)).astype(F)
    kmax = F(kk.max())
    r = F(F(rho) * F(1.0001) + F(1e-4))                         # tile_refine: the ball the test is run on
    kept_sizes = []
    for b in range(n_balls):
        keep, _ = need_list_serial(dist[b], kk, r, kmax)
        sub = table[keep]
        assert sub.shape[0] >= 1
        kept_sizes.append(sub.shape[0])
        v = rng.normal(size=(per_ball, 3))
        v *= (rng.uniform(0.0, 1.0, size=(per_ball, 1)) ** (1.0 / 3.0)) * rho / np.linalg.norm(v, axis=1, keepdims=True)
        v[0] = 0.0
        v[1:7] = np.concatenate([np.eye(3), -np.eye(3)]) * rho   # the ball's axis poles: the stencil points of empirical_normal
        pts = (c[b].astype(np.float64) + v).astype(F)
        full = orc.sdf(pts)
        part = oracle_mod.Oracle(sub).sdf(pts)
        assert np.array_equal(full.view(np.uint32), part.view(np.uint32)), f"ball {b}: kept list of {sub.shape[0]} differs from the whole table"
    assert np.mean(kept_sizes) < 40, "the rule should cull most of the primitives"


@pytest.mark.parametrize("scene_name", ["many256", "sd_obj", "sphere_box"])
def test_a_nan_point_folds_to_the_start_value_whatever_is_folded(oracle_mod, scene_name):
    """tile_mask_from_point / tile_mask_from_half_points give a point with a NaN coordinate an EMPTY list: every sphere / capsule /
    box distance at it is NaN and the fold skips NaN distances, so the whole table returns the fold's start value (FLT_MAX) too."""
    table = {"many256": lambda: scenes.many_primitives(256), "sd_obj": scenes.sd_obj, "sphere_box": scenes.sphere_box}[scene_name]()
    pts = np.array([[np.nan, 0, 0], [0.3, np.nan, 1.0], [np.nan] * 3, [1.0, 2.0, np.nan], [np.nan, np.inf, 0.0]], F)
    want = np.full(pts.shape[0], np.finfo(F).max, F)
    assert np.array_equal(oracle_mod.Oracle(table).sdf(pts).view(np.uint32), want.view(np.uint32))
    for i in range(0, table.shape[0], max(1, table.shape[0] // 16)):      # ... and so does every single primitive
        assert np.array_equal(oracle_mod.Oracle(table[i:i + 1]).sdf(pts).view(np.uint32), want.view(np.uint32))
