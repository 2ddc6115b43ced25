"""The product's packed marching-cubes tables (csrc/mc_tables.inc) are value-identical to the reference's
(cuda/includes/marching_cubes_constants.cu) - checked against the reference-derived hash and the oracle's copy."""
import ctypes
import json
import pathlib
import re

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())


def parse_inc():
    text = (ROOT / "bevy-signed-distance-mesh-generation_b200" / "csrc" / "mc_tables.inc").read_text()

    def arr(name):
        body = text.split(name)[1].split("{", 1)[1].split("};")[0]
        return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", body.replace("ull", ""))]

    return arr("SDM_MC_PACKED_INIT[256]"), arr("SDM_MC_NTRI_INIT[256]"), arr("SDM_MC_EDGEMASK_INIT[256]"), arr("SDM_MC_EDGE_CORNERS_INIT[12][2]")


def test_packed_tables_decode_to_the_reference_tables(oracle_mod):
    packed, ntri, emask, corners = parse_inc()
    assert len(packed) == 256 and len(ntri) == 256 and len(emask) == 256 and len(corners) == 24
    tri = np.full((256, 16), -1, np.int32)
    for c in range(256):
        for j in range(3 * ntri[c]):
            tri[c, j] = (packed[c] >> (4 * j)) & 0xF
        assert emask[c] == sum(1 << e for e in set(tri[c][tri[c] >= 0].tolist()))
        assert packed[c] >> (12 * ntri[c]) == 0
    assert "%016x" % oracle_mod.fnv1a64(tri) == GOLD["mc_tables"]["fnv_triangle_int32"]
    assert "%016x" % oracle_mod.fnv1a64(np.asarray(corners, np.int32)) == GOLD["mc_tables"]["fnv_edge_int32"]
    # the closed-form corner formulas used inside mc_edge_corners (k_edges / k_assign_uids)
    for e in range(12):
        c0 = (0 if e == 3 else e) if e < 4 else ((4 if e == 7 else e) if e < 8 else e - 8)
        c1 = (3 if e == 3 else e + 1) if e < 4 else ((7 if e == 7 else e + 1) if e < 8 else e - 4)
        assert (c0, c1) == (corners[2 * e], corners[2 * e + 1])


def test_packed_edge_midpoint_offsets_match_the_edge_table():
    """k_edges' fast vertex keys: the mid-point of edge e in half-voxel units, two bits per edge in SDM_EDGE_DX/DY/DZ
    (csrc/sdm_kernels.cuh), must be the mean of the edge's two corners (MC_EDGE_TABLE, marching_cubes_constants.cu:3-16;
    corner offsets of compute_mesh_generation.cu:77-86)."""
    import pathlib
    import re

    src = (pathlib.Path(__file__).resolve().parent.parent / "bevy-signed-distance-mesh-generation_b200" / "csrc" / "sdm_kernels.cuh").read_text()
    words = {ax: int(re.search(r"#define SDM_EDGE_D%s (0x[0-9a-fA-F]+)u" % ax, src).group(1), 16) for ax in "XYZ"}
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    corner = lambda c: (2 if c % 4 in (1, 2) else 0, 2 if c % 4 >= 2 else 0, 2 if c >= 4 else 0)
    for e, (a, b) in enumerate(edges):
        want = tuple((x + y) // 2 for x, y in zip(corner(a), corner(b)))
        got = tuple((words[ax] >> (2 * e)) & 3 for ax in "XYZ")
        assert got == want, (e, got, want)
