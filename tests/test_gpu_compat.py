"""The reference's kernel-level ABI (f3 of SURVEY.md section 8): load compat/compute_mesh_generation.ptx through the CUDA
DRIVER API - cuModuleLoadData + cuModuleGetFunction by symbol name + cuLaunchKernel with by-value structs, i.e. what
cudarc does for the reference's Rust host (src/cuda/mod.rs:68-90, 149-177, 226-250) - and compare the two kernels'
raw outputs with the CPU oracle (which is pinned byte-for-byte to the reference's own kernels)."""
import ctypes
import pathlib

import numpy as np
import pytest

from bsdmg_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent
PTX = ROOT / "bevy-signed-distance-mesh-generation_b200" / "compat" / "compute_mesh_generation.ptx"


class Point(ctypes.Structure):
    _fields_ = [("x", ctypes.c_float), ("y", ctypes.c_float), ("z", ctypes.c_float)]


class VoxelField(ctypes.Structure):   # bindings.h:51-55, #[repr(C)]
    _fields_ = [("voxel_size", Point), ("voxels", ctypes.c_uint64), ("voxel_count", ctypes.c_uint32)]


assert ctypes.sizeof(VoxelField) == 32


class Driver:
    def __init__(self):
        self.cu = ctypes.CDLL("libcuda.so.1")
        self.ck(self.cu.cuInit(0))
        dev = ctypes.c_int()
        self.ck(self.cu.cuDeviceGet(ctypes.byref(dev), 0))
        self.ctx = ctypes.c_void_p()
        self.ck(self.cu.cuDevicePrimaryCtxRetain(ctypes.byref(self.ctx), dev))
        self.ck(self.cu.cuCtxSetCurrent(self.ctx))
        self.mod = ctypes.c_void_p()
        ptx = PTX.read_bytes() + b"\0"
        self.ck(self.cu.cuModuleLoadData(ctypes.byref(self.mod), ptx))   # = cudarc load_ptx

    @staticmethod
    def ck(rc):
        assert rc == 0, f"CUDA driver error {rc}"

    def func(self, name):
        f = ctypes.c_void_p()
        self.ck(self.cu.cuModuleGetFunction(ctypes.byref(f), self.mod, name.encode()))
        return f

    def alloc(self, nbytes):
        p = ctypes.c_uint64()
        self.ck(self.cu.cuMemAlloc_v2(ctypes.byref(p), ctypes.c_size_t(max(nbytes, 16))))
        return p

    def h2d(self, dptr, arr):
        self.ck(self.cu.cuMemcpyHtoD_v2(dptr, arr.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(arr.nbytes)))

    def d2h(self, arr, dptr):
        self.ck(self.cu.cuMemcpyDtoH_v2(arr.ctypes.data_as(ctypes.c_void_p), dptr, ctypes.c_size_t(arr.nbytes)))

    def launch(self, f, n, *params):
        argv = (ctypes.c_void_p * len(params))(*[ctypes.cast(ctypes.byref(p), ctypes.c_void_p) for p in params])
        grid = (n + 127) // 128   # (n as f32 / BLOCK_SIZE as f32).ceil(), BLOCK_SIZE = 128 (src/cuda/mod.rs:154-161)
        self.ck(self.cu.cuLaunchKernel(f, grid, 1, 1, 128, 1, 1, 0, None, argv, None))
        self.ck(self.cu.cuCtxSynchronize())


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_reference_abi_module_matches_oracle(oracle_mod):
    assert PTX.exists(), "run __graft_entry__.build()"
    text = PTX.read_text()
    assert ".visible .entry compute_refine_voxel_field_by_sdf(" in text
    assert ".visible .entry compute_surface_triangles_from_voxel_field_by_sdf(" in text
    d = Driver()
    f_refine = d.func("compute_refine_voxel_field_by_sdf")
    f_mesh = d.func("compute_surface_triangles_from_voxel_field_by_sdf")
    o = oracle_mod.Oracle(scenes.sd_obj())
    vox, vs = o.create_voxel_field()
    for level in range(3):
        n = vox.shape[0]
        d_in, d_out = d.alloc(n * 12), d.alloc(n * 8 * 12)
        d.h2d(d_in, vox)
        fin = VoxelField(Point(*map(float, vs)), d_in.value, n)
        fout = VoxelField(Point(0.0, 0.0, 0.0), d_out.value, n * 8)          # src/cuda/mod.rs:169-173
        d.launch(f_refine, n, fin, fout)
        raw = np.empty((n * 8, 3), np.float32)
        d.d2h(raw, d_out)
        assert np.array_equal(bits(raw), bits(o.refine_raw(vox, vs))), f"refine level {level}"
        vox, vs = o.refine(vox, vs)
    n = vox.shape[0]
    d_in, d_tri = d.alloc(n * 12), d.alloc(n * 5 * 72)
    d.h2d(d_in, vox)
    d.launch(f_mesh, n, VoxelField(Point(*map(float, vs)), d_in.value, n), ctypes.c_uint64(d_tri.value))
    tris = np.empty((n * 5, 18), np.float32)
    d.d2h(tris, d_tri)
    want, _ = o.mesh_raw(vox, vs)
    assert np.array_equal(bits(tris), bits(want)), "triangle soup differs"


class RenderTexture(ctypes.Structure):   # bindings.h:38-41
    _fields_ = [("size", ctypes.c_uint * 2), ("data", ctypes.c_uint64)]


def test_reference_abi_render_module(handler):
    """compat/compute_render.ptx: the reference's `compute_render` symbol with its by-value RenderTexture / GlobalsBuffer / CameraBuffer
    parameters, loaded and launched the way cudarc does (src/cuda/mod.rs:70-79, 372-399: grid = w * h / 128, block = 128).  Same bytes as
    sdm_render on the reference's scene (which tests/test_gpu_render.py pins to the reference kernel itself)."""
    from bsdmg_b200 import handler as H

    ptx = PTX.parent / "compute_render.ptx"
    assert ptx.exists(), "run __graft_entry__.build()"
    assert ".visible .entry compute_render(" in ptx.read_text()
    d = Driver.__new__(Driver)
    d.cu = ctypes.CDLL("libcuda.so.1")
    d.ck(d.cu.cuInit(0))
    dev = ctypes.c_int()
    d.ck(d.cu.cuDeviceGet(ctypes.byref(dev), 0))
    d.ctx = ctypes.c_void_p()
    d.ck(d.cu.cuDevicePrimaryCtxRetain(ctypes.byref(d.ctx), dev))
    d.ck(d.cu.cuCtxSetCurrent(d.ctx))
    d.mod = ctypes.c_void_p()
    d.ck(d.cu.cuModuleLoadData(ctypes.byref(d.mod), ptx.read_bytes() + b"\0"))
    f = d.func("compute_render")
    w, h = 256, 144
    pos, fwd = (6.0, 3.0, 7.0), np.array([-6.0, -3.0, -7.0]) / np.linalg.norm([6.0, 3.0, 7.0])
    right = np.cross(fwd, [0.0, 1.0, 0.0]); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    f32 = lambda v: [float(np.float32(x)) for x in v]
    g = H._RenderGlobals(7, 1.5, (ctypes.c_uint * 2)(w, h), (ctypes.c_float * 2)(float(w), float(h)))
    c = H._RenderCamera((ctypes.c_float * 3)(*f32(pos)), (ctypes.c_float * 3)(*f32(fwd)), (ctypes.c_float * 3)(*f32(up)), (ctypes.c_float * 3)(*f32(right)),
                        float(np.float32(0.9)))
    d_img = d.alloc(w * h * 4)
    tex = RenderTexture((ctypes.c_uint * 2)(w, h), d_img.value)
    argv = (ctypes.c_void_p * 3)(*[ctypes.cast(ctypes.byref(p), ctypes.c_void_p) for p in (tex, g, c)])
    d.ck(d.cu.cuLaunchKernel(f, (w * h) // 128, 1, 1, 128, 1, 1, 0, None, argv, None))
    d.ck(d.cu.cuCtxSynchronize())
    got = np.empty((h, w, 4), np.uint8)
    d.d2h(got, d_img)
    handler.set_scene(scenes.render_scene())
    want = handler.render(w, h, f32(pos), f32(fwd), f32(up), f32(right), float(np.float32(0.9)), tick=7, time=1.5)
    assert len(np.unique(want.reshape(-1, 4), axis=0)) > 50
    assert np.array_equal(got, want)
