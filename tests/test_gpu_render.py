"""f4, the ray-march viewer kernel: sdm_render against the reference's own compute_render (cuda/modules/compute_render.cu:21-97)
compiled by path for the GPU with the same IEEE flags (oracle/_ref/libref_render.so).  Every byte of the RGBA image must agree."""
import ctypes
import math

import numpy as np
import pytest

import bsdmg_b200
from bsdmg_b200 import handler as H
from bsdmg_b200 import scenes

pytestmark = pytest.mark.gpu


def _camera(position, target, fov):
    """position / forward / up / right the way the reference's renderer fills CameraBuffer from a Bevy transform looking at `target`."""
    p = np.asarray(position, np.float64)
    f = np.asarray(target, np.float64) - p
    f /= np.linalg.norm(f)
    r = np.cross(f, [0.0, 1.0, 0.0]); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    return [float(np.float32(x)) for x in p], [float(np.float32(x)) for x in f], [float(np.float32(x)) for x in u], [float(np.float32(x)) for x in r], float(np.float32(fov))


@pytest.fixture(scope="module")
def refrender(oracle_mod):
    if not oracle_mod.RefRender.available():
        pytest.skip("oracle/_ref/libref_render.so not built (reference not mounted at build time)")
    return oracle_mod.RefRender()


def test_render_parameter_layouts(refrender):
    assert refrender.layout() == dict(globals=ctypes.sizeof(H._RenderGlobals), camera=ctypes.sizeof(H._RenderCamera)) == dict(globals=32, camera=52)


@pytest.mark.parametrize("w,h,pos,target,fov,screen", [
    (256, 144, (6.0, 3.0, 7.0), (0.0, 0.0, 0.0), math.pi / 4, None),          # the whole object and the wire box of the domain
    (320, 192, (1.2, 0.6, 2.0), (0.3, 0.1, 0.0), 1.1, (1280.0, 720.0)),         # close-up: many collisions at grazing angles, screen != texture size
    (128, 128, (0.0, 9.0, 0.01), (0.0, 0.0, 0.0), 0.6, None),                   # straight down; rays that leave through the depth limit
    (64, 32, (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), 2.0, None),                      # camera inside the unit sphere: immediate collisions
])
def test_render_matches_reference_kernel(refrender, handler, w, h, pos, target, fov, screen):
    handler.set_scene(scenes.render_scene())
    p, f, u, r, fv = _camera(pos, target, fov)
    got = handler.render(w, h, p, f, u, r, fv, screen_size=screen, tick=3, time=0.5)
    g = H._RenderGlobals(3, 0.5, (ctypes.c_uint * 2)(w, h), (ctypes.c_float * 2)(*(screen or (float(w), float(h)))))
    c = H._RenderCamera((ctypes.c_float * 3)(*p), (ctypes.c_float * 3)(*f), (ctypes.c_float * 3)(*u), (ctypes.c_float * 3)(*r), fv)
    want = refrender.render(g, c, w, h)
    assert got.shape == want.shape == (h, w, 4)
    assert (got[..., 3] == 255).all()
    if pos != (0.0, 0.0, 0.0):
        assert len(np.unique(want.reshape(-1, 4), axis=0)) > 50      # a real picture (inside the sphere every ray collides at the eye: one colour)
    diff = (got != want).any(axis=2)
    assert not diff.any(), f"{int(diff.sum())} of {w * h} pixels differ from the reference kernel"


def test_render_of_a_culled_scene_equals_the_full_fold(handler, monkeypatch):
    """A 256-primitive scene is rendered through the per-cell primitive masks (every marching step looks its cell up by position).
    A handle created with SDM_NO_MASKS folds the whole table at every point, which is what the reference's code does: same bytes."""
    scene = scenes.many_primitives(256)
    handler.set_scene(scene)
    p, f, u, r, fv = _camera((5.0, 4.0, 6.0), (0.0, 0.0, 0.0), 0.8)
    got = handler.render(256, 160, p, f, u, r, fv)
    assert len(np.unique(got.reshape(-1, 4), axis=0)) > 50
    monkeypatch.setenv("SDM_NO_MASKS", "1")
    h = bsdmg_b200.CudaHandler(0, scene)
    try:
        want = h.render(256, 160, p, f, u, r, fv)
        full = h.remesh(5.0, 32, 2)
    finally:
        h.close()
    assert np.array_equal(got, want)
    culled = handler.remesh(5.0, 32, 2)              # and the same for a mesh, while the two handles are at hand
    assert np.array_equal(culled.indices, full.indices) and np.array_equal(culled.positions.view(np.uint32), full.positions.view(np.uint32))
