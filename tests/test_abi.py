"""The C-ABI boundary on a machine without a GPU: the library loads, exports every function include/sdfmesh.h
declares, the struct layouts are the reference's (bindings.h:43-64), the host-only entry points work, and every
compute entry point fails loudly (no CPU fallback)."""
import ctypes
import pathlib
import re

import numpy as np
import pytest

import bsdmg_b200
from bsdmg_b200 import handler as H

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "sdfmesh.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = bsdmg_b200.load_library()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libsdfmesh.so does not export {n}"
    assert sorted(H.ABI_SYMBOLS) == names, "handler.ABI_SYMBOLS is out of sync with include/sdfmesh.h"


def test_struct_layouts():
    assert ctypes.sizeof(H._Point) == 12
    assert ctypes.sizeof(H._VoxelField) == 32 and H._VoxelField.voxels.offset == 16 and H._VoxelField.voxel_count.offset == 24
    assert bsdmg_b200.scenes.PRIM_DTYPE.itemsize == 40
    assert ctypes.sizeof(H._Mesh) == 40 and ctypes.sizeof(H._Params) == 12


def test_version_and_default_scene():
    lib = bsdmg_b200.load_library()
    assert b"sm_100a" in lib.sdm_version()
    buf = np.zeros(2, bsdmg_b200.scenes.PRIM_DTYPE)
    assert lib.sdm_scene_default(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(2)) == 2
    want = bsdmg_b200.scenes.sd_obj()
    assert buf.tobytes() == want.tobytes()   # common.cu:222-226


def test_create_voxel_field_is_host_only_and_matches_reference_order(oracle_mod):
    f = bsdmg_b200.CudaHandler.create_cuda_voxel_field()
    vox, vs = oracle_mod.Oracle.create_voxel_field()
    assert np.array_equal(f.voxels.view(np.uint32), vox.view(np.uint32)) and np.array_equal(f.voxel_size, vs)
    # x outer, z inner (src/cuda/mod.rs:110-119)
    assert np.array_equal(f.voxels[1], np.float32([-2.5, -2.5, -2.5 + 5.0 / 32]))
    g = bsdmg_b200.CudaHandler.create_cuda_voxel_field(4.0, 8)
    assert len(g) == 512 and g.voxel_size[0] == np.float32(0.5)


def test_no_cpu_fallback():
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present: the failure path cannot be observed")
    with pytest.raises(bsdmg_b200.SdfMeshError) as e:
        bsdmg_b200.CudaHandler(0)
    assert e.value.code == 3   # SDM_ERR_NO_DEVICE
