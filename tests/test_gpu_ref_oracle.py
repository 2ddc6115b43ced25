"""GPU-side oracle checks (need oracle/_ref/libref_gpu.so, which is built where /root/reference is mounted and
travels to the GPU box as a binary):

1. premise check - the reference's kernels compiled for sm_100a with IEEE flags (-fmad=false, no fast-math) equal
   the host-compiled reference (and hence the CPU oracle) bit for bit on sd_obj;
2. the functor-template restatement of the two kernels reproduces the unmodified kernels byte for byte;
3. Mandelbulb: the product equals the reference's device code (same libdevice transcendentals) bit for bit.
"""
import numpy as np
import pytest

from bsdmg_b200 import scenes

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def refgpu(oracle_mod):
    if not oracle_mod.RefGpu.available():
        pytest.skip("oracle/_ref/libref_gpu.so not built (reference not mounted at build time)")
    return oracle_mod.RefGpu()


def test_ieee_gpu_equals_cpu_on_sd_obj(refgpu, oracle_mod):
    o = oracle_mod.Oracle(scenes.sd_obj())
    vox, vs = o.create_voxel_field()
    for _ in range(2):
        raw_gpu = refgpu.refine_raw(0, vox, vs)
        raw_tpl = refgpu.refine_raw(1, vox, vs)
        raw_cpu = o.refine_raw(vox, vs)
        assert np.array_equal(bits(raw_gpu), bits(raw_cpu)), "IEEE GPU refine != CPU oracle"
        assert np.array_equal(bits(raw_tpl), bits(raw_gpu)), "functor template refine != reference kernel"
        vox, vs = o.refine(vox, vs)
    tri_gpu = refgpu.mesh_raw(0, vox, vs)
    tri_tpl = refgpu.mesh_raw(1, vox, vs)
    tri_cpu, _ = o.mesh_raw(vox, vs)
    assert np.array_equal(bits(tri_gpu), bits(tri_cpu)), "IEEE GPU mesh kernel != CPU oracle"
    assert np.array_equal(bits(tri_tpl), bits(tri_gpu)), "functor template mesh != reference kernel"


def test_mandelbulb_sdf_matches_reference_device_code(refgpu, handler):
    handler.set_scene(scenes.mandelbulb())
    rng = np.random.default_rng(7)
    pts = rng.uniform(-1.0, 1.0, size=(200_000, 3)).astype(np.float32)
    got = handler.eval_sdf(pts)
    want = refgpu.sdf(2, pts)
    same = bits(got) == bits(want)
    both_nan = np.isnan(got) & np.isnan(want)
    assert np.all(same | both_nan), f"{(~(same | both_nan)).sum()} Mandelbulb SDF values differ"


def test_mandelbulb_remesh_matches_reference_device_code(refgpu, handler, oracle_mod):
    handler.set_scene(scenes.mandelbulb())
    o = oracle_mod.Oracle(scenes.mandelbulb())   # only for the scene-independent host steps (field, weld)
    vox, vs = o.create_voxel_field(5.0, 32)
    handler.field_reset(5.0, 32)
    for _ in range(2):
        vox, vs = refgpu.refine(2, vox, vs)
        n = handler.field_refine()
        assert n == vox.shape[0]
    assert np.array_equal(bits(handler.field_download()), bits(vox))
    mesh = handler.field_to_mesh()
    tris = refgpu.mesh_raw(2, vox, vs)
    soup = handler.field_triangle_soup()
    a, b = bits(soup), bits(tris)
    nan_ok = np.isnan(soup) & np.isnan(tris)
    assert np.all((a == b) | nan_ok), f"{(~((a == b) | nan_ok)).sum()} soup words differ"
    pos, nrm, idx = o.weld(tris)
    assert np.array_equal(mesh.indices, idx)
    assert np.array_equal(bits(mesh.positions), bits(pos))
