"""ctypes mirror of the reference's ``CudaHandler`` (src/cuda/mod.rs:25-346) over libsdfmesh.so.

Method names, argument meaning and error behaviour follow the Rust host:

=====================================  =========================================================
reference (src/cuda/mod.rs)            here
=====================================  =========================================================
``CudaHandler::new()`` (:49)           ``CudaHandler(device=0)``
``create_cuda_voxel_field()`` (:105)   ``CudaHandler.create_cuda_voxel_field()``
``refine_voxel_field(&mut f)`` (:124)  ``handler.refine_voxel_field(field)`` (in place)
``voxel_field_to_mesh(&f)`` (:204)     ``handler.voxel_field_to_mesh(field) -> Mesh``
=====================================  =========================================================

plus the device-resident path (``field_reset`` / ``field_refine`` / ``field_to_mesh`` / ``remesh``) that keeps every
stage in HBM.  There is no CPU fallback: if the library or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes
import os
import dataclasses
import pathlib

import numpy as np

from .scenes import PRIM_DTYPE

_HERE = pathlib.Path(__file__).resolve().parent


class SdfMeshError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"sdfmesh error {code}: {message}")
        self.code = code


class _Point(ctypes.Structure):
    _fields_ = [("x", ctypes.c_float), ("y", ctypes.c_float), ("z", ctypes.c_float)]


class _VoxelField(ctypes.Structure):  # bindings.h:51-55
    _fields_ = [("voxel_size", _Point), ("voxels", ctypes.c_void_p), ("voxel_count", ctypes.c_uint)]


class _Params(ctypes.Structure):
    _fields_ = [("bb_size", ctypes.c_float), ("init_factor", ctypes.c_uint32), ("levels", ctypes.c_uint32)]


class _Mesh(ctypes.Structure):
    _fields_ = [
        ("positions", ctypes.c_void_p), ("normals", ctypes.c_void_p), ("indices", ctypes.c_void_p),
        ("vertex_count", ctypes.c_uint32), ("triangle_count", ctypes.c_uint32),
        ("on_device", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class _Stats(ctypes.Structure):
    _fields_ = [
        ("kernel_launches", ctypes.c_uint64), ("sdf_evals", ctypes.c_uint64), ("level_counts", ctypes.c_uint32 * 16),
        ("unique_vertices", ctypes.c_uint32), ("raw_triangles", ctypes.c_uint32), ("last_gpu_ms", ctypes.c_float),
        ("escaped_vertices", ctypes.c_uint32), ("prim_evals", ctypes.c_uint64 * 6),
        ("list_fallback_tiles", ctypes.c_uint32), ("stragglers", ctypes.c_uint32), ("newton_iterations", ctypes.c_uint64),
    ]


class _ShardInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("shard_index", "shard_count", "split_level", "split_total", "voxel_begin", "voxel_end",
                                               "final_voxels", "unique_vertices", "raw_triangles")]


class _ShardWeld(ctypes.Structure):
    _fields_ = [("vertices", ctypes.c_uint32), ("triangles", ctypes.c_uint32), ("nonfinite", ctypes.c_uint32),
                ("min_x", ctypes.c_float), ("max_x", ctypes.c_float)]


class _ShardBuffers(ctypes.Structure):
    _fields_ = [("positions", ctypes.c_void_p), ("normals", ctypes.c_void_p), ("triangle_vertex_ids", ctypes.c_void_p),
                ("capacity_vertices", ctypes.c_uint32), ("capacity_triangles", ctypes.c_uint32)]


class _RenderGlobals(ctypes.Structure):   # GlobalsBuffer, bindings.h:16-21
    _fields_ = [("tick", ctypes.c_ulonglong), ("time", ctypes.c_float), ("render_texture_size", ctypes.c_uint * 2), ("render_screen_size", ctypes.c_float * 2)]


class _RenderCamera(ctypes.Structure):    # CameraBuffer, bindings.h:23-29
    _fields_ = [("position", ctypes.c_float * 3), ("forward", ctypes.c_float * 3), ("up", ctypes.c_float * 3), ("right", ctypes.c_float * 3),
                ("fov", ctypes.c_float)]


class _PeerExport(ctypes.Structure):
    _fields_ = [("handle", (ctypes.c_ubyte * 64) * 4), ("block_bytes", ctypes.c_uint64), ("cap_vertices", ctypes.c_uint32),
                ("cap_triangles", ctypes.c_uint32), ("cap_rows", ctypes.c_uint32), ("world", ctypes.c_uint32)]


class _PeerResult(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("status", "total_vertices", "total_triangles", "vertex_offset", "triangle_offset", "vertices",
                                               "triangles")] + [("gpu_ms", ctypes.c_float)]


assert ctypes.sizeof(_VoxelField) == 32 and _VoxelField.voxels.offset == 16 and _VoxelField.voxel_count.offset == 24

# every symbol include/sdfmesh.h declares (tests/test_abi.py checks the library exports each one)
ABI_SYMBOLS = [
    "sdm_create", "sdm_destroy", "sdm_last_error", "sdm_version", "sdm_scene_default", "sdm_set_scene", "sdm_eval_sdf",
    "sdm_eval_normal", "sdm_eval_project", "sdm_create_voxel_field", "sdm_voxel_field_free", "sdm_refine_voxel_field",
    "sdm_voxel_field_to_mesh", "sdm_mesh_free", "sdm_field_reset", "sdm_field_upload", "sdm_field_refine", "sdm_field_count",
    "sdm_field_download", "sdm_field_cases", "sdm_field_to_mesh", "sdm_remesh", "sdm_mesh_download", "sdm_field_triangle_soup",
    "sdm_shard_remesh", "sdm_shard_buffers", "sdm_shard_prepare_send", "sdm_shard_reserve", "sdm_shard_weld",
    "sdm_shard_local_weld", "sdm_shard_boundary_keys", "sdm_shard_key_scratch", "sdm_shard_resolve", "sdm_shard_fixup",
    "sdm_shard_welded_buffers", "sdm_shard_reserve_welded",
    "sdm_mesh_save_obj", "sdm_hash_bytes", "sdm_reserve", "sdm_peer_root_export", "sdm_peer_attach", "sdm_peer_detach", "sdm_peer_step", "sdm_peer_finish",
    "sdm_peer_download_async", "sdm_render",
    "sdm_get_stats", "sdm_set_profiling", "sdm_get_kernel_times", "sdm_debug_fetch", "sdm_selftest_math", "sdm_mesh_download_async", "sdm_mesh_download_wait",
]


def lib_path() -> pathlib.Path:
    # SDM_LIB: developer override to load another build of the SAME CUDA library (kernel-variant experiments)
    return pathlib.Path(os.environ["SDM_LIB"]) if os.environ.get("SDM_LIB") else _HERE / "libsdfmesh.so"


_lib = None


def load_library() -> ctypes.CDLL:
    """Loads libsdfmesh.so (built in-tree by ``__graft_entry__.build()`` / csrc/Makefile).  Fails loudly."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not p.exists():
            raise FileNotFoundError(f"{p} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback")
        lib = ctypes.CDLL(str(p))
        lib.sdm_last_error.restype = ctypes.c_char_p
        lib.sdm_version.restype = ctypes.c_char_p
        lib.sdm_scene_default.restype = ctypes.c_uint32
        lib.sdm_destroy.restype = None
        lib.sdm_voxel_field_free.restype = None
        lib.sdm_mesh_free.restype = None
        lib.sdm_hash_bytes.restype = ctypes.c_uint64
        lib.sdm_hash_bytes.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        _lib = lib
    return _lib


@dataclasses.dataclass
class CudaVoxelField:
    """``CudaVoxelField`` (src/cuda/mod.rs:42-47): host list of voxel min-corners + the voxel size."""

    voxels: np.ndarray  # (n, 3) float32
    voxel_size: np.ndarray  # (3,) float32

    def __len__(self) -> int:
        return int(self.voxels.shape[0])


@dataclasses.dataclass
class Mesh:
    """What ``voxel_field_to_mesh`` returns, in the layout ``obj_to_bevy_mesh`` consumes (src/renderer/mod.rs:110-128):
    positions / normals as (V, 3) float32 and a flat triangle-list index buffer as (T, 3) uint32."""

    positions: np.ndarray
    normals: np.ndarray
    indices: np.ndarray

    @property
    def vertex_count(self) -> int:
        return int(self.positions.shape[0])

    def save_obj(self, path) -> None:
        """`obj.save("generated_mesh.obj")` of the reference's `Mesh` stage (src/renderer/mod.rs:204) through the C ABI's
        sdm_mesh_save_obj (include/sdfmesh.h has the format; the `obj` crate 0.10.2 is not vendored in the reference tree, so
        the line layout is restated from its documented writer - unpinned)."""
        lib = load_library()
        pos = np.ascontiguousarray(self.positions, np.float32)
        nrm = np.ascontiguousarray(self.normals, np.float32)
        idx = np.ascontiguousarray(self.indices, np.uint32)
        m = _Mesh(pos.ctypes.data_as(ctypes.c_void_p), nrm.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p),
                  ctypes.c_uint32(pos.shape[0]), ctypes.c_uint32(idx.shape[0]), ctypes.c_int32(0), ctypes.c_int32(0))
        rc = lib.sdm_mesh_save_obj(None, ctypes.byref(m), str(path).encode())
        if rc:
            raise SdfMeshError(rc, lib.sdm_last_error().decode())

    def bevy_attributes(self) -> dict:
        """The three buffers `obj_to_bevy_mesh` hands to Bevy (src/renderer/mod.rs:110-128): ATTRIBUTE_POSITION and
        ATTRIBUTE_NORMAL as Float32x3 arrays, Indices::U32 as the flat triangle list."""
        return {"ATTRIBUTE_POSITION": np.ascontiguousarray(self.positions, np.float32), "ATTRIBUTE_NORMAL": np.ascontiguousarray(self.normals, np.float32),
                "Indices::U32": np.ascontiguousarray(self.indices, np.uint32).reshape(-1)}

    @property
    def triangle_count(self) -> int:
        return int(self.indices.shape[0])


def _params(bb_size, init_factor, levels=0):
    return _Params(ctypes.c_float(bb_size), ctypes.c_uint32(init_factor), ctypes.c_uint32(levels))


class CudaHandler:
    BLOCK_SIZE = 128  # bindings.h:7
    MESH_GENERATION_INIT_FACTOR = 32  # bindings.h:9
    MESH_GENERATION_BB_SIZE = 5.0  # bindings.h:10

    def __init__(self, device: int = 0, scene: np.ndarray | None = None):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        self._check(self._lib.sdm_create(ctypes.c_int(device), ctypes.byref(self._h)))
        self.device = device
        if scene is not None:
            self.set_scene(scene)

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise SdfMeshError(rc, self._lib.sdm_last_error().decode())

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.sdm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- scene ------------------------------------------------------------------------------------
    def set_scene(self, scene: np.ndarray) -> None:
        scene = np.ascontiguousarray(scene, dtype=PRIM_DTYPE)
        self._check(self._lib.sdm_set_scene(self._h, scene.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(scene.shape[0])))

    def eval_sdf(self, pts) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.float32)
        self._check(self._lib.sdm_eval_sdf(self._h, pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(pts.shape[0]), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def eval_normal(self, pts) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty_like(pts)
        self._check(self._lib.sdm_eval_normal(self._h, pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(pts.shape[0]), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def eval_project(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty_like(pts)
        iters = np.empty(pts.shape[0], np.uint32)
        self._check(self._lib.sdm_eval_project(self._h, pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(pts.shape[0]),
                                               out.ctypes.data_as(ctypes.c_void_p), iters.ctypes.data_as(ctypes.c_void_p)))
        return out, iters

    # -- the reference's CudaHandler surface (host buffers) -----------------------------------------
    @staticmethod
    def create_cuda_voxel_field(bb_size: float = 5.0, init_factor: int = 32) -> CudaVoxelField:
        lib = load_library()
        f = _VoxelField()
        p = _params(bb_size, init_factor)
        rc = lib.sdm_create_voxel_field(ctypes.byref(p), ctypes.byref(f))
        if rc:
            raise SdfMeshError(rc, lib.sdm_last_error().decode())
        n = int(f.voxel_count)
        vox = np.ctypeslib.as_array(ctypes.cast(f.voxels, ctypes.POINTER(ctypes.c_float)), shape=(n, 3)).copy()
        vs = np.array([f.voxel_size.x, f.voxel_size.y, f.voxel_size.z], np.float32)
        lib.sdm_voxel_field_free(ctypes.byref(f))
        return CudaVoxelField(vox, vs)

    def _upload(self, field: CudaVoxelField) -> None:
        vox = np.ascontiguousarray(field.voxels, np.float32).reshape(-1, 3)
        vs = np.asarray(field.voxel_size, np.float32)
        f = _VoxelField(_Point(float(vs[0]), float(vs[1]), float(vs[2])), vox.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(vox.shape[0]))
        self._check(self._lib.sdm_field_upload(self._h, ctypes.byref(f)))

    def refine_voxel_field(self, field: CudaVoxelField) -> None:
        """In place, like the reference: list replaced by the surviving children, size halved; empty = no-op (:137)."""
        if len(field) == 0:
            return
        self._upload(field)
        n = self.field_refine()
        field.voxels = self.field_download(n)
        field.voxel_size = (np.asarray(field.voxel_size, np.float32) / np.float32(2.0)).astype(np.float32)

    def voxel_field_to_mesh(self, field: CudaVoxelField) -> Mesh:
        if len(field) == 0:  # :327-345
            return Mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
        self._upload(field)
        return self.field_to_mesh()

    # -- device-resident path -----------------------------------------------------------------------
    def field_reset(self, bb_size: float = 5.0, init_factor: int = 32) -> None:
        p = _params(bb_size, init_factor)
        self._check(self._lib.sdm_field_reset(self._h, ctypes.byref(p)))

    def field_upload(self, field: CudaVoxelField) -> None:
        self._upload(field)

    def field_refine(self) -> int:
        n = ctypes.c_uint32(0)
        self._check(self._lib.sdm_field_refine(self._h, ctypes.byref(n)))
        return int(n.value)

    def field_count(self):
        n = ctypes.c_uint32(0)
        vs = _Point()
        self._check(self._lib.sdm_field_count(self._h, ctypes.byref(n), ctypes.byref(vs)))
        return int(n.value), np.array([vs.x, vs.y, vs.z], np.float32)

    def field_download(self, n: int | None = None) -> np.ndarray:
        if n is None:
            n, _ = self.field_count()
        out = np.empty((n, 3), np.float32)
        self._check(self._lib.sdm_field_download(self._h, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(n)))
        return out

    def field_cases(self) -> np.ndarray:
        n, _ = self.field_count()
        out = np.empty(n, np.uint8)
        self._check(self._lib.sdm_field_cases(self._h, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(n)))
        return out

    def field_triangle_soup(self) -> np.ndarray:
        """Reference raw format: (5*n, 18) float32 = 5 Triangle slots per voxel (bindings.h:57-64)."""
        n, _ = self.field_count()
        out = np.empty((5 * n, 18), np.float32)
        self._check(self._lib.sdm_field_triangle_soup(self._h, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(5 * n)))
        return out

    def _download(self, m: _Mesh) -> Mesh:
        pos = np.empty((m.vertex_count, 3), np.float32)
        nrm = np.empty((m.vertex_count, 3), np.float32)
        idx = np.empty((m.triangle_count, 3), np.uint32)
        self._check(self._lib.sdm_mesh_download(self._h, ctypes.byref(m), pos.ctypes.data_as(ctypes.c_void_p),
                                                nrm.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p)))
        return Mesh(pos, nrm, idx)

    def field_to_mesh(self, download: bool = True):
        m = _Mesh()
        self._check(self._lib.sdm_field_to_mesh(self._h, ctypes.byref(m)))
        return self._download(m) if download else m

    def remesh(self, bb_size: float = 5.0, init_factor: int = 32, levels: int = 0, download: bool = True):
        """Level-0 field, ``levels`` refinements and the mesh, all on the device.  With ``download=False`` returns the
        raw ``SdmMesh`` view (device pointers valid until the next mesh call)."""
        p = _params(bb_size, init_factor, levels)
        m = _Mesh()
        self._check(self._lib.sdm_remesh(self._h, ctypes.byref(p), ctypes.byref(m)))
        return self._download(m) if download else m

    # -- shards (multi-GPU) -------------------------------------------------------------------------
    def shard_remesh(self, bb_size, init_factor, levels, split_level, shard_index, shard_count) -> dict:
        p = _params(bb_size, init_factor, levels)
        info = _ShardInfo()
        self._check(self._lib.sdm_shard_remesh(self._h, ctypes.byref(p), ctypes.c_uint32(split_level), ctypes.c_uint32(shard_index),
                                               ctypes.c_uint32(shard_count), ctypes.byref(info)))
        return {n: int(getattr(info, n)) for n, _ in _ShardInfo._fields_}

    def shard_buffers(self) -> dict:
        b = _ShardBuffers()
        self._check(self._lib.sdm_shard_buffers(self._h, ctypes.byref(b)))
        return dict(positions=int(b.positions or 0), normals=int(b.normals or 0), triangle_vertex_ids=int(b.triangle_vertex_ids or 0),
                    capacity_vertices=int(b.capacity_vertices), capacity_triangles=int(b.capacity_triangles))

    def shard_prepare_send(self, vertex_offset: int) -> None:
        self._check(self._lib.sdm_shard_prepare_send(self._h, ctypes.c_uint32(vertex_offset)))

    def shard_reserve(self, total_vertices: int, total_triangles: int) -> None:
        self._check(self._lib.sdm_shard_reserve(self._h, ctypes.c_uint32(total_vertices), ctypes.c_uint32(total_triangles)))

    def shard_weld(self, total_vertices: int, total_triangles: int, download: bool = False):
        m = _Mesh()
        self._check(self._lib.sdm_shard_weld(self._h, ctypes.c_uint32(total_vertices), ctypes.c_uint32(total_triangles), ctypes.byref(m)))
        return self._download(m) if download else m

    # distributed weld: local weld, boundary keys resolved on rank 0, concatenation (include/sdfmesh.h)
    def shard_local_weld(self) -> dict:
        w = _ShardWeld()
        self._check(self._lib.sdm_shard_local_weld(self._h, ctypes.byref(w)))
        return {n: getattr(w, n) for n, _ in _ShardWeld._fields_}

    def shard_boundary_keys(self, intervals):
        """intervals: [(lo, hi)] closed x intervals (the other shards' ranges) -> (device pointer to rows of 4 u32, row count)."""
        n = len(intervals)
        lo = (ctypes.c_float * max(n, 1))(*[a for a, _ in intervals])
        hi = (ctypes.c_float * max(n, 1))(*[b for _, b in intervals])
        ptr, cnt = ctypes.c_void_p(), ctypes.c_uint32(0)
        self._check(self._lib.sdm_shard_boundary_keys(self._h, lo, hi, ctypes.c_uint32(n), ctypes.byref(ptr), ctypes.byref(cnt)))
        return int(ptr.value or 0), int(cnt.value)

    def shard_key_scratch(self, rows: int) -> int:
        ptr = ctypes.c_void_p()
        self._check(self._lib.sdm_shard_key_scratch(self._h, ctypes.c_uint32(rows), ctypes.byref(ptr)))
        return int(ptr.value or 0)

    def shard_resolve(self, rows_ptr: int, total_rows: int, vertex_counts) -> list:
        """rank 0: resolves the gathered key rows; returns the number of removed (duplicate) vertices per shard."""
        n = len(vertex_counts)
        vc = (ctypes.c_uint32 * n)(*vertex_counts)
        removed = (ctypes.c_uint32 * n)()
        self._check(self._lib.sdm_shard_resolve(self._h, ctypes.c_void_p(rows_ptr), ctypes.c_uint32(total_rows), vc, ctypes.c_uint32(n), removed))
        return list(removed)

    def shard_fixup(self, triangle_counts, download: bool = False):
        """rank 0: after the welded shards have arrived at their concatenated offsets - drops duplicates, makes indices global."""
        tc = (ctypes.c_uint32 * len(triangle_counts))(*triangle_counts)
        m = _Mesh()
        self._check(self._lib.sdm_shard_fixup(self._h, tc, ctypes.byref(m)))
        return self._download(m) if download else m

    def shard_welded_buffers(self) -> dict:
        b = _ShardBuffers()
        self._check(self._lib.sdm_shard_welded_buffers(self._h, ctypes.byref(b)))
        return dict(positions=int(b.positions or 0), normals=int(b.normals or 0), indices=int(b.triangle_vertex_ids or 0))

    def shard_reserve_welded(self, total_vertices: int, total_triangles: int) -> None:
        self._check(self._lib.sdm_shard_reserve_welded(self._h, ctypes.c_uint32(total_vertices), ctypes.c_uint32(total_triangles)))

    def save_obj(self, m, path) -> None:
        """sdm_mesh_save_obj of a device-resident mesh (`remesh(..., download=False)`)."""
        self._check(self._lib.sdm_mesh_save_obj(self._h, ctypes.byref(m), str(path).encode()))

    # -- ray-march viewer -----------------------------------------------------------------------------------------------------
    def render(self, width: int, height: int, position, forward, up, right, fov: float, screen_size=None, tick: int = 0, time: float = 0.0) -> np.ndarray:
        """``CudaHandler::render`` (src/cuda/mod.rs:348-409): (height, width, 4) uint8 image of the handle's current scene."""
        g = _RenderGlobals(tick, time, (ctypes.c_uint * 2)(width, height), (ctypes.c_float * 2)(*(screen_size or (float(width), float(height)))))
        c = _RenderCamera((ctypes.c_float * 3)(*position), (ctypes.c_float * 3)(*forward), (ctypes.c_float * 3)(*up), (ctypes.c_float * 3)(*right), fov)
        out = np.empty((height, width, 4), np.uint8)
        self._check(self._lib.sdm_render(self._h, ctypes.byref(g), ctypes.byref(c), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    # -- peer exchange (include/sdfmesh.h): the distributed weld driven from the device -------------------------------
    def reserve(self, voxel_capacity: int) -> None:
        self._check(self._lib.sdm_reserve(self._h, ctypes.c_uint32(voxel_capacity)))

    def peer_root_export(self, world: int, cap_rows: int) -> bytes:
        e = _PeerExport()
        self._check(self._lib.sdm_peer_root_export(self._h, ctypes.c_uint32(world), ctypes.c_uint32(cap_rows), ctypes.byref(e)))
        return bytes(e)

    def peer_attach(self, export: bytes, rank: int, world: int, same_process_root: "CudaHandler | None" = None) -> None:
        e = _PeerExport.from_buffer_copy(export)
        root = same_process_root._h if same_process_root is not None else None
        self._check(self._lib.sdm_peer_attach(self._h, ctypes.byref(e), ctypes.c_uint32(rank), ctypes.c_uint32(world), root))

    def peer_detach(self) -> None:
        self._check(self._lib.sdm_peer_detach(self._h))

    def peer_step(self, bb_size, init_factor, levels, split_level, epoch, deliver=0, phase_mask=31, spin=True) -> None:
        p = _params(bb_size, init_factor, levels)
        self._check(self._lib.sdm_peer_step(self._h, ctypes.byref(p), ctypes.c_uint32(split_level), ctypes.c_uint32(epoch), ctypes.c_int(deliver),
                                            ctypes.c_uint32(phase_mask), ctypes.c_int(1 if spin else 0)))

    def peer_finish(self):
        """-> (result dict, SdmMesh view): rank 0 with deliver = 0 gets the merged mesh, deliver = 1 gives every rank its own rows."""
        r, m = _PeerResult(), _Mesh()
        self._check(self._lib.sdm_peer_finish(self._h, ctypes.byref(r), ctypes.byref(m)))
        return {n: getattr(r, n) for n, _ in _PeerResult._fields_}, m

    def peer_download_async(self, result: dict, positions_ptr: int, normals_ptr: int, indices_ptr: int) -> None:
        r = _PeerResult(**result)
        self._check(self._lib.sdm_peer_download_async(self._h, ctypes.byref(r), ctypes.c_void_p(positions_ptr), ctypes.c_void_p(normals_ptr),
                                                      ctypes.c_void_p(indices_ptr)))

    def sync(self) -> None:
        """Waits for everything enqueued on the handle's stream (used between the phases of an emulated multi-rank step)."""
        n = ctypes.c_uint32(0)
        self._check(self._lib.sdm_field_count(self._h, ctypes.byref(n), None))

    def download_into(self, m, positions_ptr: int, normals_ptr: int, indices_ptr: int) -> None:
        """sdm_mesh_download into caller-provided (e.g. pinned) host memory."""
        self._check(self._lib.sdm_mesh_download(self._h, ctypes.byref(m), ctypes.c_void_p(positions_ptr), ctypes.c_void_p(normals_ptr),
                                                ctypes.c_void_p(indices_ptr)))

    def debug_fetch(self, name: str, count: int, dtype=np.float32) -> np.ndarray:
        out = np.empty(count, dtype)
        self._check(self._lib.sdm_debug_fetch(self._h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(out.nbytes)))
        return out

    def selftest_math(self, div_samples: int = 1 << 32) -> dict:
        out = (ctypes.c_ulonglong * 4)()
        self._check(self._lib.sdm_selftest_math(self._h, ctypes.c_ulonglong(div_samples), out))
        return dict(sqrt_mismatches=int(out[0]), sqrt_slow_path=int(out[1]), div_mismatches=int(out[2]), div_slow_path=int(out[3]))

    def download_into_async(self, m, positions_ptr: int, normals_ptr: int, indices_ptr: int) -> None:
        """sdm_mesh_download_async: the copy overlaps the next remesh (outputs are double-buffered in the handle)."""
        self._check(self._lib.sdm_mesh_download_async(self._h, ctypes.byref(m), ctypes.c_void_p(positions_ptr), ctypes.c_void_p(normals_ptr),
                                                      ctypes.c_void_p(indices_ptr)))

    def download_wait(self) -> None:
        self._check(self._lib.sdm_mesh_download_wait(self._h))

    def set_profiling(self, enabled: bool) -> None:
        self._check(self._lib.sdm_set_profiling(self._h, ctypes.c_int(1 if enabled else 0)))

    def kernel_times(self):
        """[(kernel name, ms)] of the last remesh (needs set_profiling(True))."""
        names = (ctypes.c_char_p * 64)()
        ms = (ctypes.c_float * 64)()
        n = self._lib.sdm_get_kernel_times(self._h, names, ms, ctypes.c_uint32(64))
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def stats(self) -> dict:
        s = _Stats()
        self._check(self._lib.sdm_get_stats(self._h, ctypes.byref(s)))
        return dict(kernel_launches=int(s.kernel_launches), sdf_evals=int(s.sdf_evals), level_counts=list(s.level_counts),
                    unique_vertices=int(s.unique_vertices), raw_triangles=int(s.raw_triangles), last_gpu_ms=float(s.last_gpu_ms),
                    escaped_vertices=int(s.escaped_vertices), list_fallback_tiles=int(s.list_fallback_tiles), stragglers=int(s.stragglers),
                    newton_iterations=int(s.newton_iterations),
                    prim_evals=dict(zip(("refine", "classify", "project", "tail", "normals", "orient"), [int(x) for x in s.prim_evals])))
