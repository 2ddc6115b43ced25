"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracle (oracle/liborc.so) and, when it has been
built in the container that holds /root/reference, for the host-compiled reference (oracle/_ref/libref_host.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
_F = ctypes.POINTER(ctypes.c_float)
_U32 = ctypes.POINTER(ctypes.c_uint32)
_U8 = ctypes.POINTER(ctypes.c_uint8)


def _fp(a):
    return a.ctypes.data_as(_F)


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref where the reference is mounted) with oracle/Makefile."""
    if force:
        subprocess.run(["make", "-C", str(HERE), "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", str(HERE), "all"], check=True, capture_output=True)


def fnv1a64(data) -> int:
    """FNV-1a 64 over raw bytes (vectorised: hashes of golden arrays up to a few hundred MB)."""
    b = np.frombuffer(memoryview(data).cast("B"), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).ravel()
    h = 0xCBF29CE484222325
    prime = 0x100000001B3
    mask = 0xFFFFFFFFFFFFFFFF
    # plain loop in chunks through python ints is too slow for MBs; use the C helper when available
    lib = _load_orc(optional=True)
    if lib is not None and hasattr(lib, "orc_fnv1a64"):
        lib.orc_fnv1a64.restype = ctypes.c_uint64
        return int(lib.orc_fnv1a64(b.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(b.size)))
    for x in b.tobytes():
        h = ((h ^ x) * prime) & mask
    return h


_orc = None


def _load_orc(optional: bool = False):
    global _orc
    if _orc is None:
        path = HERE / "liborc.so"
        if not path.exists():
            try:
                build()
            except Exception:
                if optional:
                    return None
                raise
        _orc = ctypes.CDLL(str(path))
    return _orc


class Oracle:
    """numpy front-end of oracle/sdm_oracle.cpp (scene = SdmPrimitive structured array)."""

    def __init__(self, scene: np.ndarray):
        self.lib = _load_orc()
        self.scene = np.ascontiguousarray(scene)
        self._sp = self.scene.ctypes.data_as(ctypes.c_void_p)
        self._sn = ctypes.c_uint32(self.scene.shape[0])

    @staticmethod
    def threads() -> int:
        return int(_load_orc().orc_num_threads())

    @staticmethod
    def set_threads(n: int) -> None:
        _load_orc().orc_set_num_threads(int(n))

    def sdf(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.float32)
        self.lib.orc_eval_sdf(self._sp, self._sn, _fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(out))
        return out

    def normal(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty_like(pts)
        self.lib.orc_eval_normal(self._sp, self._sn, _fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(out))
        return out

    def project(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty_like(pts)
        iters = np.empty(pts.shape[0], np.uint32)
        self.lib.orc_eval_project(self._sp, self._sn, _fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(out), iters.ctypes.data_as(_U32))
        return out, iters

    @staticmethod
    def create_voxel_field(bb_size: float = 5.0, init_factor: int = 32):
        vox = np.empty((init_factor ** 3, 3), np.float32)
        vs = np.empty(3, np.float32)
        _load_orc().orc_create_voxel_field(ctypes.c_float(bb_size), ctypes.c_uint32(init_factor), _fp(vox), _fp(vs))
        return vox, vs

    def refine_raw(self, vox, vs):
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        vs = np.ascontiguousarray(vs, np.float32)
        out = np.empty((vox.shape[0] * 8, 3), np.float32)
        self.lib.orc_refine_raw(self._sp, self._sn, _fp(vox), ctypes.c_uint32(vox.shape[0]), _fp(vs), _fp(out))
        return out

    def refine(self, vox, vs):
        """refine_voxel_field (src/cuda/mod.rs:124-202): kernel + stable retain + halved size."""
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        if vox.shape[0] == 0:
            return vox, np.asarray(vs, np.float32) / np.float32(2.0) * 0 + np.asarray(vs, np.float32)  # :137 no-op
        raw = self.refine_raw(vox, vs)
        n = int(self.lib.orc_retain_finite(_fp(raw), ctypes.c_uint32(raw.shape[0])))
        return raw[:n].copy(), (np.asarray(vs, np.float32) / np.float32(2.0)).astype(np.float32)

    def mesh_raw(self, vox, vs):
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        vs = np.ascontiguousarray(vs, np.float32)
        n = vox.shape[0]
        tris = np.empty((n * 5, 18), np.float32)
        cases = np.empty(n, np.uint8)
        self.lib.orc_mesh_raw(self._sp, self._sn, _fp(vox), ctypes.c_uint32(n), _fp(vs), _fp(tris), cases.ctypes.data_as(_U8))
        return tris, cases

    @staticmethod
    def weld(tris):
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 18)
        n = tris.shape[0]
        pos = np.empty((3 * n, 3), np.float32)
        nrm = np.empty((3 * n, 3), np.float32)
        idx = np.empty((n, 3), np.uint32)
        nv, nt = ctypes.c_uint32(0), ctypes.c_uint32(0)
        _load_orc().orc_weld(_fp(tris), ctypes.c_uint32(n), _fp(pos), _fp(nrm), idx.ctypes.data_as(_U32), ctypes.byref(nv), ctypes.byref(nt))
        return pos[: nv.value].copy(), nrm[: nv.value].copy(), idx[: nt.value].copy()

    def mesh(self, vox, vs):
        """voxel_field_to_mesh (src/cuda/mod.rs:204-346): positions, normals, indices, cases."""
        tris, cases = self.mesh_raw(vox, vs)
        pos, nrm, idx = self.weld(tris)
        return pos, nrm, idx, cases

    def remesh(self, bb_size=5.0, init_factor=32, levels=0):
        vox, vs = self.create_voxel_field(bb_size, init_factor)
        counts = [vox.shape[0]]
        for _ in range(levels):
            vox, vs = self.refine(vox, vs)
            counts.append(vox.shape[0])
        pos, nrm, idx, cases = self.mesh(vox, vs)
        return dict(voxels=vox, voxel_size=vs, level_counts=counts, positions=pos, normals=nrm, indices=idx, cases=cases)


class RefHost:
    """The reference's own kernels, host-compiled (oracle/_ref/libref_host.so).  Scene is always sd_obj."""

    def __init__(self):
        path = HERE / "_ref" / "libref_host.so"
        if not path.exists():
            raise FileNotFoundError(str(path))
        self.lib = ctypes.CDLL(str(path))
        self.lib.ref_bb_size.restype = ctypes.c_float
        self.lib.ref_smooth_min.restype = ctypes.c_float
        self.lib.ref_smooth_min.argtypes = [ctypes.c_float] * 3

    @staticmethod
    def available() -> bool:
        return (HERE / "_ref" / "libref_host.so").exists()

    def refine_raw(self, vox, vs):
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        vs = np.ascontiguousarray(vs, np.float32)
        out = np.empty((vox.shape[0] * 8, 3), np.float32)
        self.lib.ref_refine(_fp(vox), ctypes.c_uint32(vox.shape[0]), _fp(vs), _fp(out))
        return out

    def mesh_raw(self, vox, vs):
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        vs = np.ascontiguousarray(vs, np.float32)
        tris = np.empty((vox.shape[0] * 5, 18), np.float32)
        self.lib.ref_mesh(_fp(vox), ctypes.c_uint32(vox.shape[0]), _fp(vs), _fp(tris))
        return tris

    # ---- functor templates over the reference's own functions (ref_functor.inc): scene 1 = sd_obj, 2 = sd_unit_mandelbulb,
    #      3 = SdmPrimitive table fold (un-culled) --------------------------------------------------------------------------
    def set_threads(self, n: int) -> None:
        self.lib.ref_set_num_threads(int(n))

    def threads(self) -> int:
        return int(self.lib.ref_num_threads())

    def tpl_refine_raw(self, scene_id, table, vox, vs):
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        vs = np.ascontiguousarray(vs, np.float32)
        table = np.ascontiguousarray(table) if table is not None else np.zeros(0, np.uint8)
        out = np.empty((vox.shape[0] * 8, 3), np.float32)
        self.lib.ref_tpl_refine(ctypes.c_int(scene_id), table.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(table.shape[0]), _fp(vox),
                                ctypes.c_uint32(vox.shape[0]), _fp(vs), _fp(out))
        return out

    def tpl_mesh_raw(self, scene_id, table, vox, vs):
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        vs = np.ascontiguousarray(vs, np.float32)
        table = np.ascontiguousarray(table) if table is not None else np.zeros(0, np.uint8)
        tris = np.empty((vox.shape[0] * 5, 18), np.float32)
        self.lib.ref_tpl_mesh(ctypes.c_int(scene_id), table.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(table.shape[0]), _fp(vox),
                              ctypes.c_uint32(vox.shape[0]), _fp(vs), _fp(tris))
        return tris

    def tpl_sdf(self, table, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        table = np.ascontiguousarray(table)
        out = np.empty(pts.shape[0], np.float32)
        self.lib.ref_tpl_sdf(table.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(table.shape[0]), _fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(out))
        return out

    def _pts(self, fn, pts, width):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty((pts.shape[0], width) if width > 1 else pts.shape[0], np.float32)
        getattr(self.lib, fn)(_fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(out))
        return out

    def sd_obj(self, pts):
        return self._pts("ref_sd_obj", pts, 1)

    def normal_sd_obj(self, pts):
        return self._pts("ref_empirical_normal_sd_obj", pts, 3)

    def project_sd_obj(self, pts):
        return self._pts("ref_closest_surface_point_sd_obj", pts, 3)

    def sd_unit_mandelbulb(self, pts):
        return self._pts("ref_sd_unit_mandelbulb", pts, 1)

    def sd_unit_sphere(self, pts):
        return self._pts("ref_sd_unit_sphere", pts, 1)

    def sd_box(self, pts, bp, bs):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.float32)
        self.lib.ref_sd_box(_fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(np.asarray(bp, np.float32)), _fp(np.asarray(bs, np.float32)), _fp(out))
        return out

    def sd_line(self, pts, b0, b1):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.float32)
        self.lib.ref_sd_line(_fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(np.asarray(b0, np.float32)), _fp(np.asarray(b1, np.float32)), _fp(out))
        return out

    def sd_box_skeleton(self, pts, bp, bs, lw):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.float32)
        self.lib.ref_sd_box_skeleton(_fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(np.asarray(bp, np.float32)), _fp(np.asarray(bs, np.float32)), ctypes.c_float(lw), _fp(out))
        return out

    def smooth_min(self, a, b, k):
        return np.float32(self.lib.ref_smooth_min(a, b, k))

    def mc_tables(self):
        e = np.empty(24, np.int32)
        t = np.empty(4096, np.int32)
        self.lib.ref_mc_tables(e.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p))
        return e, t

    def layout(self):
        L = self.lib
        return dict(point=L.ref_sizeof_point(), voxel_field=L.ref_sizeof_voxel_field(), voxels_at=L.ref_offsetof_voxels(),
                    count_at=L.ref_offsetof_voxel_count(), vertex=L.ref_sizeof_vertex(), triangle=L.ref_sizeof_triangle(),
                    block_size=L.ref_block_size(), init_factor=L.ref_init_factor(), bb_size=float(L.ref_bb_size()))


class RefGpu:
    """The reference's kernels compiled by nvcc for sm_100a with IEEE flags (oracle/_ref/libref_gpu.so).
    scene: 0 = unmodified reference kernels (sd_obj), 1 = functor templates with sd_obj, 2 = with sd_unit_mandelbulb,
    3 = with the SdmPrimitive table fold of set_table() (reference primitives, un-culled)."""

    def __init__(self):
        path = HERE / "_ref" / "libref_gpu.so"
        if not path.exists():
            raise FileNotFoundError(str(path))
        self.lib = ctypes.CDLL(str(path))

    @staticmethod
    def available() -> bool:
        return (HERE / "_ref" / "libref_gpu.so").exists()

    def set_table(self, table) -> None:
        table = np.ascontiguousarray(table)
        rc = self.lib.refgpu_set_table(table.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(table.shape[0]))
        assert rc == 0

    def refine_raw(self, scene, vox, vs):
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        vs = np.ascontiguousarray(vs, np.float32)
        out = np.empty((vox.shape[0] * 8, 3), np.float32)
        rc = self.lib.refgpu_refine(ctypes.c_int(scene), _fp(vox), ctypes.c_uint32(vox.shape[0]), _fp(vs), _fp(out))
        assert rc == 0
        return out

    def refine(self, scene, vox, vs):
        raw = self.refine_raw(scene, vox, vs)
        keep = np.isfinite(raw).all(axis=1)
        return raw[keep].copy(), (np.asarray(vs, np.float32) / np.float32(2.0)).astype(np.float32)

    def mesh_raw(self, scene, vox, vs):
        vox = np.ascontiguousarray(vox, np.float32).reshape(-1, 3)
        vs = np.ascontiguousarray(vs, np.float32)
        tris = np.empty((vox.shape[0] * 5, 18), np.float32)
        rc = self.lib.refgpu_mesh(ctypes.c_int(scene), _fp(vox), ctypes.c_uint32(vox.shape[0]), _fp(vs), _fp(tris))
        assert rc == 0
        return tris

    def sdf(self, scene, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.float32)
        rc = self.lib.refgpu_sdf(ctypes.c_int(scene), _fp(pts), ctypes.c_uint32(pts.shape[0]), _fp(out))
        assert rc == 0
        return out


class RefRender:
    """The reference's ray-march module (cuda/modules/compute_render.cu) compiled by path for sm_100a with IEEE flags
    (oracle/_ref/libref_render.so)."""

    def __init__(self):
        path = HERE / "_ref" / "libref_render.so"
        if not path.exists():
            raise FileNotFoundError(str(path))
        self.lib = ctypes.CDLL(str(path))

    @staticmethod
    def available() -> bool:
        return (HERE / "_ref" / "libref_render.so").exists()

    def layout(self):
        return dict(globals=int(self.lib.refrender_sizeof_globals()), camera=int(self.lib.refrender_sizeof_camera()))

    def render(self, globals_struct, camera_struct, width, height):
        out = np.empty((height, width, 4), np.uint8)
        rc = self.lib.refrender(ctypes.byref(globals_struct), ctypes.byref(camera_struct), out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        return out
