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


def test_obj_writer_layout(tmp_path):
    """Mesh.save_obj: the layout the reference's `obj.save` produces for its ObjData (src/cuda/mod.rs:303-326)."""
    m = bsdmg_b200.Mesh(np.float32([[0, 0.5, -1.25], [1, 1e-5, 2], [3, 4, 5]]), np.float32([[0, 0, 1], [0, 1, 0], [1, 0, 0]]), np.uint32([[0, 1, 2]]))
    p = tmp_path / "generated_mesh.obj"
    m.save_obj(p)
    lines = p.read_text().splitlines()
    assert lines[0] == "v 0 0.5 -1.25" and lines[1] == "v 1 0.00001 2"
    assert lines[3] == "vt 0 0" and lines[4] == "vn 0 0 1"
    assert lines[7:9] == ["o default", "g default"] and lines[9] == "f 1/1/1 2/1/2 3/1/3"
    a = m.bevy_attributes()
    assert a["Indices::U32"].tolist() == [0, 1, 2] and a["ATTRIBUTE_POSITION"].shape == (3, 3)


def test_header_is_plain_c_and_layouts_match_the_ctypes_mirrors(tmp_path):
    """include/sdfmesh.h is the drop-in boundary: it must compile as C99 on its own, and every struct the Python mirror
    (handler.py) passes across it must have the size the C compiler gives it."""
    import subprocess

    structs = {"SdmPoint": H._Point, "SdmVoxelField": H._VoxelField, "SdmParams": H._Params, "SdmMesh": H._Mesh, "SdmStats": H._Stats,
               "SdmShardInfo": H._ShardInfo, "SdmShardBuffers": H._ShardBuffers, "SdmShardWeld": H._ShardWeld,
               "SdmPeerExport": H._PeerExport, "SdmPeerResult": H._PeerResult}
    src = tmp_path / "layout.c"
    body = "\n".join(f'    printf("{n} %zu\\n", sizeof({n}));' for n in structs)
    src.write_text('#include <stdio.h>\n#include "sdfmesh.h"\nint main(void) {\n' + body +
                   '\n    printf("SdmPrimitive %zu\\n", sizeof(SdmPrimitive));\n    printf("SdmTriangle %zu\\n", sizeof(SdmTriangle));\n    return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for n, c in structs.items():
        assert int(out[n]) == ctypes.sizeof(c), f"{n}: C says {out[n]} bytes, ctypes mirror has {ctypes.sizeof(c)}"
    assert int(out["SdmPrimitive"]) == bsdmg_b200.scenes.PRIM_DTYPE.itemsize == 40
    assert int(out["SdmTriangle"]) == 72 and int(out["SdmPoint"]) == 12 and int(out["SdmVoxelField"]) == 32   # bindings.h:43-64


def _rust_f32(x):
    """Rust's `{}` for an f32, restated with numpy's Dragon4 (shortest digits that round-trip, positional)."""
    x = np.float32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    return np.format_float_positional(x, unique=True, trim="-")


def test_obj_writer_numbers_and_empty_mesh(tmp_path):
    """sdm_mesh_save_obj (C ABI, host mesh - no GPU needed): every number is the shortest round-tripping decimal in
    positional notation (Rust's `{}`), over random bit patterns, powers of two, denormals, +-0, NaN and infinities; an
    empty mesh writes the header lines only (src/cuda/mod.rs:327-345)."""
    rng = np.random.default_rng(11)
    vals = np.concatenate([
        rng.integers(0, 1 << 32, size=30000, dtype=np.uint64).astype(np.uint32).view(np.float32),
        rng.uniform(-2.5, 2.5, size=30000).astype(np.float32),
        np.float32(2.0) ** np.arange(-149, 128, dtype=np.float32),
        np.float32([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-5, 0.1, 1 / 3, 16777216.0, 3.4028235e38, 1e-45, 1.17549435e-38]),
    ]).astype(np.float32)
    vals = np.resize(vals, (len(vals) + 2) // 3 * 3).reshape(-1, 3)
    m = bsdmg_b200.Mesh(vals, vals[::-1].copy(), np.uint32([[0, 1, 2], [2, 1, 0]]))
    p = tmp_path / "numbers.obj"
    m.save_obj(p)
    lines = p.read_text().splitlines()
    nv = vals.shape[0]
    assert len(lines) == 2 * nv + 1 + 2 + 2
    for i in range(nv):
        assert lines[i] == "v " + " ".join(_rust_f32(x) for x in vals[i]), (i, vals[i])
        got = np.array([np.float32(t) for t in lines[i].split()[1:]], np.float32)   # and they read back as the same floats
        assert np.array_equal(got.view(np.uint32)[~np.isnan(vals[i])], vals[i].view(np.uint32)[~np.isnan(vals[i])])
    assert lines[nv] == "vt 0 0"
    assert lines[nv + 1] == "vn " + " ".join(_rust_f32(x) for x in vals[-1])
    assert lines[-2:] == ["f 1/1/1 2/1/2 3/1/3", "f 3/1/3 2/1/2 1/1/1"]
    e = bsdmg_b200.Mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    e.save_obj(tmp_path / "empty.obj")
    assert (tmp_path / "empty.obj").read_text() == "vt 0 0\no default\ng default\n"


def test_hash_bytes_is_fnv1a64(oracle_mod):
    lib = bsdmg_b200.load_library()
    data = np.random.default_rng(5).integers(0, 256, size=100_003, dtype=np.uint8)
    assert int(lib.sdm_hash_bytes(data.ctypes.data, data.size)) == oracle_mod.fnv1a64(data)
    assert int(lib.sdm_hash_bytes(None, 0)) == 0xCBF29CE484222325
