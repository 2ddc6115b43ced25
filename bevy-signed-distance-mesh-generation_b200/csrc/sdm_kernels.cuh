// sdm_kernels.cuh - the sm_100a kernels of the mesh-generation path.
//
//   k_init_field      level-0 dense list                       (src/cuda/mod.rs:105-122)
//   k_refine          3x3x3 lattice signs per parent           (compute_mesh_generation.cu:12-62)
//   k_refine_emit     stable compaction of surviving children, their case indices (:51, src/cuda/mod.rs:179-194)
//   k_cases           8 corner signs -> case index (compute_mesh_generation.cu:77-86, marching_cubes.cu:19-23); only when the
//                     last k_refine's lattice signs cannot be reused
//   k_tri_offsets     triangle offsets per voxel from the case table (marching_cubes.cu:24-25)
//   k_edges           edge mid-points (marching_cubes.cu:13-16) de-duplicated by exact bit pattern in a hash table
//                     (fast 64-bit lattice keys or generic float-bit keys); vertex ids, start points, list records
//   k_project(+_tail) closest_surface_point per distinct mid-point (signed_distance.cu:227-240)
//   k_build_masks(_fine)  per-cell primitive masks for large scenes (exact culling of the fold), zero-crossing flags
//   k_vertex_normals  empirical_normal per projected vertex (signed_distance.cu:181-202) + the vertex's weld key
//   k_orient          per triangle: face normal vs. centroid normal, flip (compute_mesh_generation.cu:103-113),
//                     finite filter (src/cuda/mod.rs:289), first-occurrence slot per vertex
//   k_weld_keys / k_weld_min / k_weld_mark / k_bitscan / k_emit_*   the reference-order weld (src/cuda/mod.rs:263-296)
//   k_soup            the reference's raw 5-slot Triangle format (compute_mesh_generation.cu:107-118)
//   k_shard_* / k_res_* / k_fix_*   multi-GPU: shard selection, host-driven distributed weld
//   k_peer_*          multi-GPU: the same weld driven from the device over peer-mapped memory (flags, key rows, pairs, pushes)
//
// All kernels read their problem sizes from DevState in device memory, so that a whole remesh is enqueued without any host
// synchronisation.  A kernel that evaluates the SDF never waits for another block: every ordered step (compaction, offsets,
// ids, ranks) is a streaming kernel of its own with a block-granular decoupled look-back.
#pragma once

#include <cooperative_groups.h>

#include "sdm_device.cuh"
#include "mc_tables.inc"

namespace sdm {

namespace cg = cooperative_groups;

enum Ticket : int { TK_REFINE0 = 0, TK_CLASSIFY = 16, TK_PROJECT = 17, TK_SCAN_FIRST = 18, TK_SCAN_TRI = 19, TK_TAIL = 20, TK_ASSIGN = 21, TK_REFINE_EMIT = 22, TK_NORMALS = 23, TK_ORIENT = 24, TK_EDGES = 25, TK_COUNT = 26 };
enum ErrFlag : uint32_t {
    ERR_VOXEL_CAP = 1u, ERR_TRI_CAP = 2u, ERR_UNIQ_CAP = 4u, ERR_HASH_FULL = 8u,
    ERR_PEER_TIMEOUT = 32u,   // peer exchange: a flag from another rank did not arrive within 20 s
    ERR_LATTICE = 16u   // a voxel is not on the integer lattice the fast vertex keys assume: the host retries with the generic keys
};

struct DevState {
    uint32_t level_count[17];   // active voxels per level (index = level)
    uint32_t n_tris_raw;        // triangles before the finite filter
    uint32_t n_uniq;            // distinct edge mid-points
    uint32_t n_tris_out;        // triangles after the finite filter
    uint32_t n_verts_out;       // welded vertices
    uint32_t error_flags;
    uint32_t ticket[TK_COUNT];
    uint32_t n_stragglers;      // vertices handed from k_project to k_project_tail
    uint32_t cases_from_refine; // epoch of the last k_refine that met an inexact lattice (its case indices must not be used)
    uint32_t weld_dups;         // vertices whose quantised weld key was already in the table (0: the weld merges nothing)
    uint32_t peer_scan_total[2];   // totals of the peer exchange's bitmap scans (unused by the host)
    uint32_t n_escaped;         // vertices whose Newton iterate left the region their inherited list is proven for (general path)
    uint32_t list_fallbacks;    // tiles of the mesh stage that had to use the cell masks instead of inherited lists
    uint32_t n_orient_pending;  // triangles whose orientation the six-sample test of k_orient<true> left open (k_orient_pending)
    unsigned long long prim_evals[6];   // (primitive, point) distance evaluations: refine, classify, project, tail, normals, orient
    unsigned long long newton_iters;   // total closest_surface_point iterations (statistics)
    unsigned long long tail_steps, tail_rebuilds, tail_unlisted;   // k_project_tail: warp steps, list rebuilds, steps without a tile list (sdm_debug_fetch "state")
};

// Marching-cubes tables staged per block
struct McShared {
    unsigned long long packed[256];
    unsigned short edgemask[256];
    unsigned char ntri[256];
};
__constant__ unsigned long long c_mc_packed[256];
__constant__ unsigned short c_mc_edgemask[256];
__constant__ unsigned char c_mc_ntri[256];

__device__ __forceinline__ void stage_mc(McShared* s) {
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        s->packed[i] = c_mc_packed[i];
        s->edgemask[i] = c_mc_edgemask[i];
        s->ntri[i] = c_mc_ntri[i];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_init_field(float* __restrict__ vox, DevState* st, float bb_size, uint32_t init, float size,
                                                    uint32_t cap_vox) {
    const uint64_t n = (uint64_t) init * init * init;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->level_count[0] = (uint32_t) (n <= cap_vox ? n : 0);
        if (n > cap_vox) atomicOr(&st->error_flags, ERR_VOXEL_CAP);
    }
    if (n > cap_vox) return;
    const float half = bb_size / 2.0f;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t z = (uint32_t) (i % init), y = (uint32_t) ((i / init) % init), x = (uint32_t) (i / ((uint64_t) init * init));
        vox[3 * i + 0] = (float) x * size - half;   // src/cuda/mod.rs:114-116
        vox[3 * i + 1] = (float) y * size - half;
        vox[3 * i + 2] = (float) z * size - half;
    }
}

// primitives one evaluation folds in the current tile (for the work counters)
__device__ __forceinline__ uint32_t tile_prims(const SceneView& sc) {
    if (!sc.wmask) return sc.nprims;
    const uint32_t tc = *sc.tcount;
    if (tc != SDM_TLIST_NONE) return tc;
    uint32_t c = 0;
    for (uint32_t w = 0; w < sc.W; w++) c += __popc(sc.wmask[w]);
    return c;
}
enum WorkKind : int { WK_REFINE = 0, WK_CLASSIFY = 1, WK_PROJECT = 2, WK_TAIL = 3, WK_NORMALS = 4, WK_ORIENT = 5 };

__device__ __forceinline__ uint32_t warp_inclusive_sum(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t) o) v += t;
    }
    return v;
}

// Refinement in two kernels.
// k_refine: one warp = one tile of 32 parents, one lane = one parent.  The reference tests the 8 corners of each of the 8
// children (<= 64 evaluations, compute_mesh_generation.cu:30-49); the corners are the 27 points of the parent's 3x3x3
// lattice (`upper` of child i is `lower` of child i+1: the same float expression base + vec3{i,j,k} * size), so 27
// evaluations give the same 64 signs.  They are done as three x-slabs of 9 points held in registers (ILP 9 per lane).  The
// kernel only records the 27 signs per parent; tiles need no order, so no warp ever waits for another one (with the ordered
// append fused in, a third of the kernel's instructions were look-back spins: ncu, profiles/).
// k_refine_emit: surviving children are appended in the reference's order (n_id = id*8 + i*4 + j*2 + k, :51; Vec::retain
// is stable, src/cuda/mod.rs:192) at the offset given by a block scan + decoupled look-back across tiles - a streaming pass.
#ifndef SDM_REFINE_MINB
#define SDM_REFINE_MINB 3
#endif
__global__ void __launch_bounds__(256, SDM_REFINE_MINB) k_refine(const uint4* __restrict__ scene, const float* __restrict__ in_vox, DevState* st, int level,
                                                float osx, float osy, float osz, MaskGrid grid, uint32_t* __restrict__ out_m27,
                                                uint32_t cases_epoch /* 0: no case indices wanted */, int use_cell_flags,
                                                const uint4* __restrict__ vl_in /* records of the previous level's parents, or null */,
                                                const uint32_t* __restrict__ vparent_in /* record index per voxel of this level */,
                                                uint4* __restrict__ vl_out /* one record per voxel of this level, or null */, float delta) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = st->level_count[level];
    const uint32_t ntiles = (n + 31u) >> 5;
    const bool want_cases = cases_epoch != 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->ticket[TK_REFINE_EMIT] = 0;   // k_refine_emit runs after this kernel
    unsigned long long work = 0;
    bool lattice_ok = true;
    while (true) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(&st->ticket[TK_REFINE0 + level], 1u);   // dynamic hand-out: tile costs vary with the list lengths
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) {
            if (lane == 0 && work) atomicAdd(&st->prim_evals[WK_REFINE], work);
            if (want_cases && !__all_sync(0xffffffffu, lattice_ok) && lane == 0) st->cases_from_refine = cases_epoch;
            break;
        }
        const uint32_t p0 = tile << 5;
        const uint32_t np = min(32u, n - p0);
        const bool active = lane < np;
        const uint32_t p = p0 + lane;   // this lane's parent
        float bx = 0.f, by = 0.f, bz = 0.f;
        if (active) {
            bx = in_vox[3 * (size_t) p + 0];
            by = in_vox[3 * (size_t) p + 1];
            bz = in_vox[3 * (size_t) p + 2];
        }
        // Dense level of a culled scene: a parent inside cells that provably contain no zero crossing has 27 equal signs, so
        // none of its children survives (is_border, :36-49) - it is not evaluated and does not lengthen the tile's list.
        const bool eval = active && (!use_cell_flags || box_may_cross(grid, bx, by, bz, bx + 2.0f * osx, by + 2.0f * osy, bz + 2.0f * osz));
        const uint32_t neval = (uint32_t) __popc(__ballot_sync(0xffffffffu, eval));
        // Primitives that can matter anywhere inside this tile's parent voxels (edge = 2 * child size) inflated by delta: the
        // lane's own need-list on that region is recorded (vl_out) and inherited by its children (and by the mesh stage, whose
        // evaluation points - Newton iterates, stencil points, centroids - may leave the voxel by up to delta).  Candidates: the
        // union of the records the lanes inherited from THEIR parents, else the cell masks.
        if (neval && sc.wmask) {
            // tile list: the lanes' parent boxes as they are (the 27 lattice points lie inside); own record: the same boxes inflated
            // by delta (+ the empirical_normal stencil reach of the mesh stage, 2e-3) - one pass over the candidates decides both
            const float lx = bx, ly = by, lz = bz;
            const float hx = bx + 2.0f * osx, hy = by + 2.0f * osy, hz = bz + 2.0f * osz;
            uint16_t* own = vl_out ? reinterpret_cast<uint16_t*>(vl_out + 2 * (size_t) p) : nullptr;   // warp-uniform null / non-null
            uint32_t ncand = SDM_TLIST_NONE;
            if (vl_in) {
                uint4 lo = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), hi = lo;
                if (eval) {
                    const uint32_t pi = vparent_in[p];
                    lo = __ldg(vl_in + 2 * (size_t) pi); hi = __ldg(vl_in + 2 * (size_t) pi + 1);
                }
                ncand = tile_union_lists(sc, lo, hi);
            }
            if (ncand != SDM_TLIST_NONE) tile_refine(sc, ncand, eval, lx, ly, lz, hx, hy, hz, 0.0f, own, delta, 0.0021f);
            else tile_mask_from_box(grid, sc, eval, lx, ly, lz, hx, hy, hz, own, delta, 0.0021f);
        }
        if (neval) work += (unsigned long long) tile_prims(sc) * 27u * neval;
        uint32_t m27 = 0;
        if (eval) {
#pragma unroll 1
            for (int a = 0; a < 3; a++) {
                float px[9], py[9], pz[9], f[9];
#pragma unroll
                for (int q = 0; q < 9; q++) {
                    // base + vec3{i,j,k} * output_voxel_size (compute_mesh_generation.cu:33-34)
                    px[q] = bx + (float) a * osx;
                    py[q] = by + (float) (q / 3) * osy;
                    pz[q] = bz + (float) (q % 3) * osz;
                }
                eval_scene<9>(sc, px, py, pz, f);
#pragma unroll
                for (int q = 0; q < 9; q++) m27 |= (uint32_t) (f[q] <= 0.0f) << (a * 9 + q);   // obj_contains, :8-10
            }
        }
        if (active) {
            out_m27[p] = m27;   // a skipped parent: 27 equal signs, recorded as "all outside" (no child survives either way)
            if (want_cases) {
                // The mesh stage samples child corners at child_base + size (compute_mesh_generation.cu:77-86); the lattice
                // has base + 2*size where the child is the upper one.  The 27 signs double as the children's corner signs
                // only if both expressions give the same float on every axis (always true on the dyadic default grid).
                lattice_ok = lattice_ok && __float_as_uint((bx + osx) + osx) == __float_as_uint(bx + 2.0f * osx) &&
                             __float_as_uint((by + osy) + osy) == __float_as_uint(by + 2.0f * osy) &&
                             __float_as_uint((bz + osz) + osz) == __float_as_uint(bz + 2.0f * osz);
            }
        }
    }
}

__device__ __forceinline__ uint32_t refine_keep_mask(uint32_t m27) {
    uint32_t keep = 0;
#pragma unroll
    for (int ch = 0; ch < 8; ch++) {
        const uint32_t M = child_mask(ch >> 2, (ch >> 1) & 1, ch & 1);
        const uint32_t sgn = m27 & M;
        keep |= (uint32_t) (sgn != 0u && sgn != M) << ch;   // is_border: corners do not all agree (:36-49)
    }
    return keep;
}
// one parent per thread: a warp's children land in one contiguous run, which the warp writes together
__global__ void __launch_bounds__(256) k_refine_emit(const float* __restrict__ in_vox, float* __restrict__ out_vox, DevState* st, int level,
                                                     uint32_t epoch, uint64_t* tiles, uint32_t cap_vox, float osx, float osy, float osz,
                                                     const uint32_t* __restrict__ in_m27, uint8_t* __restrict__ out_cases,
                                                     uint32_t* __restrict__ vparent_out /* record index (= parent) per child, or null */) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_w[10];
    __shared__ uint8_t s_desc[8][256];   // per warp: (owner lane << 3 | child) of every output slot of the warp's run
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = st->level_count[level];
    const uint32_t ntiles = (n + blockDim.x - 1u) / blockDim.x;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&st->ticket[TK_REFINE_EMIT], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) {
            if (tile == 0 && threadIdx.x == 0) st->level_count[level + 1] = 0;   // empty input: no-op (src/cuda/mod.rs:137)
            break;
        }
        const uint32_t p = tile * blockDim.x + threadIdx.x;
        const bool active = p < n;
        const uint32_t m27 = active ? in_m27[p] : 0u;
        uint32_t keep = 0;
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
            const uint32_t M = child_mask(ch >> 2, (ch >> 1) & 1, ch & 1);
            const uint32_t sgn = m27 & M;
            keep |= (uint32_t) (sgn != 0u && sgn != M) << ch;   // is_border: corners do not all agree (:36-49)
        }
        const uint32_t cnt = __popc(keep);
        const uint32_t incl = warp_inclusive_sum(cnt, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t block_end;
        const uint32_t base = block_lookback(tiles, tile, epoch, total, s_w, block_end);
        if (block_end > cap_vox) {
            if (threadIdx.x == 0) atomicOr(&st->error_flags, ERR_VOXEL_CAP);
        } else if (total) {
            // A warp's children are one contiguous run of the output, [base, base + total).  Written thread by thread (each parent its
            // own <= 8 children) every store instruction touched up to 32 sectors; here the run is written by the warp: a descriptor
            // (owner lane, child) per output slot in shared memory, then 32 slots at a time - the parent's base through a shuffle,
            // the same float expression for the child, and the 96 floats of 32 children as three fully coalesced stores.
            uint8_t* desc = s_desc[threadIdx.x >> 5];
            {
                uint32_t o = incl - cnt;
#pragma unroll
                for (int ch = 0; ch < 8; ch++)
                    if (keep & (1u << ch)) desc[o++] = (uint8_t) ((lane << 3) | (uint32_t) ch);
            }
            __syncwarp();
            float bx = 0.f, by = 0.f, bz = 0.f;
            if (keep) { bx = in_vox[3 * (size_t) p + 0]; by = in_vox[3 * (size_t) p + 1]; bz = in_vox[3 * (size_t) p + 2]; }
            for (uint32_t k0 = 0; k0 < total; k0 += 32u) {
                const uint32_t k = k0 + lane;
                const bool valid = k < total;
                const uint32_t d = valid ? desc[k] : 0u, owner = d >> 3;
                const int ch = (int) (d & 7u);
                const float obx = __shfl_sync(0xffffffffu, bx, owner), oby = __shfl_sync(0xffffffffu, by, owner), obz = __shfl_sync(0xffffffffu, bz, owner);
                const uint32_t om27 = __shfl_sync(0xffffffffu, m27, owner);
                const float x = obx + (float) (ch >> 2) * osx, y = oby + (float) ((ch >> 1) & 1) * osy, z = obz + (float) (ch & 1) * osz;
                float* run = out_vox + 3 * (size_t) (base + k0);
#pragma unroll
                for (uint32_t i = 0; i < 3u; i++) {
                    const uint32_t f = 32u * i + lane, c = f / 3u, comp = f - 3u * c;
                    const float vx = __shfl_sync(0xffffffffu, x, c), vy = __shfl_sync(0xffffffffu, y, c), vz = __shfl_sync(0xffffffffu, z, c);
                    if (k0 + c < total) run[f] = comp == 0u ? vx : (comp == 1u ? vy : vz);
                }
                if (valid) {
                    if (out_cases) {
                        // cube_index bit c = sign at corner c: +x iff c%4 in {1,2}, +y iff c%4 >= 2, +z iff c >= 4
                        uint32_t cube = 0;
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const int dx = ((c & 3) == 1 || (c & 3) == 2) ? 1 : 0, dy = ((c & 3) >= 2) ? 1 : 0, dz = (c >= 4) ? 1 : 0;
                            const int l = ((ch >> 2) + dx) * 9 + (((ch >> 1) & 1) + dy) * 3 + ((ch & 1) + dz);
                            cube |= ((om27 >> l) & 1u) << c;
                        }
                        out_cases[base + k] = (uint8_t) cube;
                    }
                    if (vparent_out) vparent_out[base + k] = p - lane + owner;
                }
            }
            __syncwarp();   // the descriptors are rewritten by the next tile
        }
        if (tile == ntiles - 1 && threadIdx.x == 0) st->level_count[level + 1] = min(block_end, cap_vox);
    }
}

// ---- classification and edge vertices -------------------------------------------------------------------------
// The stage is split into four small kernels so that the hash phase (latency-bound) runs without any barrier or look-back:
//   k_cases        8 corner evaluations -> cube_index (marching_cubes.cu:19-23); skipped on the device when the last
//                  k_refine already wrote the case indices from its lattice signs
//   k_tri_offsets  triangle count per voxel from the case table, exclusive scan in list order -> tri_off, n_tris_raw
//   k_edges        every edge the case uses gets its mid-point mix(a, b, 0.5f) (marching_cubes.cu:13-16), de-duplicated
//                  across voxels by the exact start point in a hash table (see "Fast vertex keys" below): an identical
//                  start point gives an identical projection, normal and weld key, so it is projected once instead of
//                  once per incident triangle (~6x); mid-points that differ in any bit stay separate and are merged - if
//                  at all - by the reference's quantised weld later, exactly as the reference would.
//                  slot_ref[3*triangle + corner] = table entry of that corner's vertex; the vertices a tile of 256 voxels
//                  created get consecutive ids, their start points and the list record of the creating voxel.
__device__ __forceinline__ void mc_edge_corners(int e, int& c0, int& c1) {   // MC_EDGE_TABLE (marching_cubes_constants.cu:3-16)
    c0 = (e < 4) ? ((e == 3) ? 0 : e) : (e < 8 ? ((e == 7) ? 4 : e) : e - 8);
    c1 = (e < 4) ? ((e == 3) ? 3 : e + 1) : (e < 8 ? ((e == 7) ? 7 : e + 1) : e - 4);
}
// mid-point of edge e of the voxel at (bx,by,bz): mix(a, b, 0.5f) = a * (1.0f - 0.5f) + b * 0.5f with the corners of
// compute_mesh_generation.cu:77-86
__device__ __forceinline__ void edge_midpoint(float bx, float by, float bz, float sx, float sy, float sz, int e, float& mx, float& my, float& mz) {
    int c0, c1;
    mc_edge_corners(e, c0, c1);
    float ax, ay, az, cx, cy, cz;
    voxel_corner(bx, by, bz, sx, sy, sz, c0, ax, ay, az);
    voxel_corner(bx, by, bz, sx, sy, sz, c1, cx, cy, cz);
    mx = ax * (1.0f - 0.5f) + cx * 0.5f; my = ay * (1.0f - 0.5f) + cy * 0.5f; mz = az * (1.0f - 0.5f) + cz * 0.5f;
}

__global__ void __launch_bounds__(256) k_cases(const uint4* __restrict__ scene, const float* __restrict__ vox, DevState* st, int level,
                                               uint8_t* __restrict__ cases, float sx, float sy, float sz, MaskGrid grid, uint32_t cases_epoch) {
    extern __shared__ uint4 smem[];
    if (cases_epoch != 0u && st->cases_from_refine != cases_epoch) return;   // the last k_refine's lattice signs are the case indices
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = st->level_count[level];
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5, warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long work = 0;
    for (uint32_t v0 = warp_id << 5; v0 < n; v0 += warps_total << 5) {
        const uint32_t v = v0 + lane;
        const bool active = v < n;
        float bx = 0.f, by = 0.f, bz = 0.f;
        if (active) { bx = vox[3 * (size_t) v]; by = vox[3 * (size_t) v + 1]; bz = vox[3 * (size_t) v + 2]; }
        float cxs[8], cys[8], czs[8];
#pragma unroll
        for (int c = 0; c < 8; c++) voxel_corner(bx, by, bz, sx, sy, sz, c, cxs[c], cys[c], czs[c]);   // :77-86
        tile_mask_from_box(grid, sc, active, bx, by, bz, bx + sx, by + sy, bz + sz);
        work += (unsigned long long) tile_prims(sc) * 8u * min(32u, n - v0);
        if (active) {
            float f[8];
            eval_scene<8>(sc, cxs, cys, czs, f);
            uint32_t cube_index = 0;
#pragma unroll
            for (int c = 0; c < 8; c++) cube_index |= (uint32_t) (f[c] <= 0.0f) << c;   // marching_cubes.cu:22
            cases[v] = (uint8_t) cube_index;
        }
    }
    if (lane == 0 && work) atomicAdd(&st->prim_evals[WK_CLASSIFY], work);
}

// tri_off[v] = number of triangles of the voxels before v (list order); 4 voxels per thread, block-granular look-back
__global__ void __launch_bounds__(256) k_tri_offsets(DevState* st, int level, const uint8_t* __restrict__ cases, uint32_t* __restrict__ tri_off,
                                                     uint32_t epoch, uint64_t* tiles, uint32_t cap_tris) {
    __shared__ unsigned char s_ntri[256];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_w[10];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = c_mc_ntri[i];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = st->level_count[level];
    const uint32_t per_tile = blockDim.x * 4u;
    const uint32_t ntiles = (n + per_tile - 1u) / per_tile;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&st->ticket[TK_CLASSIFY], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) {
            if (tile == 0 && threadIdx.x == 0) st->n_tris_raw = 0;
            break;
        }
        const uint32_t v0 = tile * per_tile + threadIdx.x * 4u;
        uint32_t c[4] = { 0, 0, 0, 0 };
        if (v0 + 3u < n) {
            const uchar4 q = *reinterpret_cast<const uchar4*>(cases + v0);
            c[0] = s_ntri[q.x]; c[1] = s_ntri[q.y]; c[2] = s_ntri[q.z]; c[3] = s_ntri[q.w];
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (v0 + j < n) c[j] = s_ntri[cases[v0 + j]];
        }
        const uint32_t mine = c[0] + c[1] + c[2] + c[3];
        const uint32_t incl = warp_inclusive_sum(mine, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t end;
        const uint32_t base = block_lookback(tiles, tile, epoch, total, s_w, end);
        uint32_t o = base + incl - mine;
        if (v0 + 3u < n) {
            *reinterpret_cast<uint4*>(tri_off + v0) = make_uint4(o, o + c[0], o + c[0] + c[1], o + c[0] + c[1] + c[2]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) { if (v0 + j < n) tri_off[v0 + j] = o; o += c[j]; }
        }
        if (tile == ntiles - 1 && threadIdx.x == 0) {
            if (end > cap_tris) atomicOr(&st->error_flags, ERR_TRI_CAP);
            st->n_tris_raw = min(end, cap_tris);
        }
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifndef SDM_PREFETCH
#define SDM_PREFETCH 1   /* next chunk / tile on its way to L1 while the current one is evaluated: k_project 3.25 -> 3.13 ms on configs[2] */
#endif

// Fast vertex keys.  On a dyadic grid (the default 5 / 2^k one, any grid whose coordinates are exact multiples of half a voxel)
// an edge mid-point is identified by three integers, its coordinates in units of half a voxel, so a table entry is ONE 64-bit
// word (key + 1; 0 = empty) claimed with a 64-bit CAS - half the table of the generic path (16-byte entries keyed by the float
// bits), which moved 1.5 GB through DRAM for 135 MB of keys (ncu, profiles/).  A tile first de-duplicates its keys in SHARED
// memory (an edge is used by up to four voxels, most of them neighbours in the list and hence in the tile), so only one thread
// per distinct key of the tile goes to the global table: ~2.5x fewer global probes, and almost no CAS contention.
// (Placing neighbouring keys next to each other in the global table - one 512-byte bucket per 4^3 voxels - was tried and is 4x
// SLOWER: a tile then hammers a handful of L2 lines instead of spreading over all slices; profiles/r2a_*.)
// Exactness: the kernel CHECKS, for every voxel and axis, that base, base + size and their mid-point are bit-equal to
// o + n * half for the integers n it uses; equal keys then imply bit-equal mid-points, which is all the de-duplication needs
// (two equal mid-points with different keys would merely be projected twice and merged by the weld, like in the reference).
// A voxel that fails the check raises ERR_LATTICE and the host repeats the mesh stage with the generic keys.
struct EdgeLattice { float ox, oy, oz; float hx, hy, hz; float ihx, ihy, ihz; uint32_t enabled; };

__device__ __forceinline__ bool lattice_axis(float b, float s, float o, float h, float ih, int& n) {
    n = __float2int_rn((b - o) * ih);
    const float fn = (float) n;
    const float hi = b + s;
    const float mid = b * (1.0f - 0.5f) + hi * 0.5f;
    return n >= 0 && n < (1 << 21) - 2 && __float_as_uint(o + fn * h) == __float_as_uint(b) &&
           __float_as_uint(o + (fn + 2.0f) * h) == __float_as_uint(hi) && __float_as_uint(o + (fn + 1.0f) * h) == __float_as_uint(mid);
}
// lattice coordinates (half-voxel units, relative to the voxel's base index) of the mid-point of edge e: 0, 1 or 2 per axis, two bits
// per edge packed into one word per axis (edge e joins the corners of MC_EDGE_TABLE[e]; corner c: +x iff c%4 in {1,2}, +y iff c%4 >= 2,
// +z iff c >= 4; tests/test_tables.py checks the packed words against the tables)
#define SDM_EDGE_DX 0x281919u
#define SDM_EDGE_DY 0xa06464u
#define SDM_EDGE_DZ 0x55aa00u
__device__ __forceinline__ void edge_lattice_offset(int e, int& dx, int& dy, int& dz) {
    dx = (int) ((SDM_EDGE_DX >> (2 * e)) & 3u); dy = (int) ((SDM_EDGE_DY >> (2 * e)) & 3u); dz = (int) ((SDM_EDGE_DZ >> (2 * e)) & 3u);
}
__device__ __forceinline__ unsigned long long lattice_key(int ux, int uy, int uz) {
    return 1ull + ((unsigned long long) (uint32_t) ux | ((unsigned long long) (uint32_t) uy << 21) | ((unsigned long long) (uint32_t) uz << 42));
}
__device__ __forceinline__ uint32_t hash_key64(unsigned long long k) { return hash96((uint32_t) k, (uint32_t) (k >> 32), 0x9E3779B9u); }
// find-or-insert of a 64-bit key (never 0) from `pos`; returns the entry, *won = this call created it; 0xFFFFFFFF = table full
__device__ __forceinline__ uint32_t hash64_probe_from(unsigned long long* table, uint32_t mask, uint32_t pos, unsigned long long key, bool* won) {
    for (uint32_t probe = 0; probe <= mask; probe++) {
        unsigned long long cur = __ldcg(table + pos);
        if (cur == 0ull) {
            cur = atomicCAS(table + pos, 0ull, key);
            if (cur == 0ull) { *won = true; return pos; }
        }
        if (cur == key) { *won = false; return pos; }
        pos = (pos + 1) & mask;
    }
    *won = false;
    return 0xFFFFFFFFu;
}

// Edge vertices.  One tile = 256 consecutive voxels, handed out by ticket (tiles close in the list run close in time).  Every edge
// the case uses is looked up / claimed in the vertex table; the vertices a tile created get consecutive ids from ONE atomicAdd
// on n_uniq per tile - ids are consecutive within a tile and tiles are nearly in list order, which is all the later per-vertex
// kernels need (coherent neighbourhoods for their primitive lists, coalesced loads).  The id numbering is internal: output
// order comes from the weld (first occurrence in triangle order), not from here.  For every created vertex: its start point
// (mid-point, marching_cubes.cu:13-16), its list record (the creating voxel's parent) and the id in the entry's side array.
#define SDM_EDGE_SLOTS 2048u   /* shared-memory key slots per tile (256 voxels x ~4 edges, ~2.5 uses per key) */
template <bool LATTICE>
__global__ void __launch_bounds__(256) k_edges(const float* __restrict__ vox, DevState* st, int level, const uint8_t* __restrict__ cases,
                                               const uint32_t* __restrict__ tri_off, void* table_raw, uint32_t table_mask,
                                               uint32_t* __restrict__ slot_ref, uint32_t* __restrict__ tri_rec, const uint32_t* __restrict__ vparent,
                                               uint32_t* __restrict__ entry_uid, float* __restrict__ ustart, uint32_t* __restrict__ urec,
                                               EdgeLattice lat, float sx, float sy, float sz, uint32_t cap_uniq) {
    __shared__ McShared mc;
    __shared__ uint32_t s_eref[12 * 256];   // [edge][thread]: bank-conflict-free dynamic indexing by edge
    __shared__ unsigned long long s_key[LATTICE ? SDM_EDGE_SLOTS : 1];   // the tile's distinct keys
    __shared__ uint32_t s_gpos[LATTICE ? SDM_EDGE_SLOTS : 1];            // their global entries (bit 31: this tile created the entry)
    __shared__ uint32_t s_w[9];
    __shared__ uint32_t s_tile, s_base;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        mc.packed[i] = c_mc_packed[i]; mc.edgemask[i] = c_mc_edgemask[i]; mc.ntri[i] = c_mc_ntri[i];
    }
    __syncthreads();
    if (st->error_flags) return;
    const uint32_t n = st->level_count[level];
    const uint32_t ntiles = (n + 255u) >> 8;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4* table16 = reinterpret_cast<uint4*>(table_raw);
    unsigned long long* table8 = reinterpret_cast<unsigned long long*>(table_raw);
    bool full = false, off_lattice = false;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&st->ticket[TK_EDGES], 1u);
        if (LATTICE)
            for (uint32_t i = threadIdx.x; i < SDM_EDGE_SLOTS; i += blockDim.x) s_key[i] = 0ull;
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) break;
        const uint32_t v = (tile << 8) + threadIdx.x;
        const uint32_t cube_index = v < n ? cases[v] : 0u;
        const uint32_t emask = mc.edgemask[cube_index];   // case 0 uses no edge
        uint32_t won_mask = 0;
        float bx = 0.f, by = 0.f, bz = 0.f;
        if (LATTICE) {
            // phase 1: the voxel's keys into the tile's shared-memory table; own_mask = keys this thread put there first.
            // A key that finds the shared table full goes to the global table directly (s_eref bit 31).
            uint32_t own_mask = 0;
            int nx = 0, ny = 0, nz = 0;
            if (emask) {
                bx = vox[3 * (size_t) v]; by = vox[3 * (size_t) v + 1]; bz = vox[3 * (size_t) v + 2];
                const bool ok = lattice_axis(bx, sx, lat.ox, lat.hx, lat.ihx, nx) & lattice_axis(by, sy, lat.oy, lat.hy, lat.ihy, ny) &
                                lattice_axis(bz, sz, lat.oz, lat.hz, lat.ihz, nz);
                if (!ok) { off_lattice = true; nx = ny = nz = 0; }
                for (uint32_t m = emask; m; m &= m - 1u) {
                    const int e = __ffs((int) m) - 1;
                    int dx, dy, dz;
                    edge_lattice_offset(e, dx, dy, dz);
                    const unsigned long long key = lattice_key(nx + dx, ny + dy, nz + dz);
                    uint32_t p = hash_key64(key) >> 8 & (SDM_EDGE_SLOTS - 1u), ref = 0xFFFFFFFFu;
                    for (uint32_t probe = 0; probe < 64u; probe++) {
                        unsigned long long cur = s_key[p];
                        if (cur == 0ull) {
                            cur = atomicCAS(&s_key[p], 0ull, key);
                            if (cur == 0ull) {   // this thread owns the key in the tile: it will do the global find-or-insert
                                own_mask |= 1u << e; ref = p;
                                const uint32_t g0 = hash_key64(key) & table_mask;
                                s_gpos[p] = g0;
                                prefetch_l2(table8 + g0);
                                break;
                            }
                        }
                        if (cur == key) { ref = p; break; }
                        p = (p + 1u) & (SDM_EDGE_SLOTS - 1u);
                    }
                    if (ref == 0xFFFFFFFFu) {   // crowded shared table (pathological tile): straight to the global table
                        bool w;
                        const uint32_t gp = hash64_probe_from(table8, table_mask, hash_key64(key) & table_mask, key, &w);
                        if (gp == 0xFFFFFFFFu) { full = true; w = false; }
                        if (w) won_mask |= 1u << e;
                        ref = 0x80000000u | (gp & 0x7FFFFFFFu);
                    }
                    s_eref[e * 256 + threadIdx.x] = ref;
                }
                // phase 2: one global find-or-insert per distinct key of the tile (its table line is already on its way)
                for (uint32_t m = own_mask; m; m &= m - 1u) {
                    const int e = __ffs((int) m) - 1;
                    const uint32_t p = s_eref[e * 256 + threadIdx.x];
                    const unsigned long long key = s_key[p];
                    bool w;
                    const uint32_t gp = hash64_probe_from(table8, table_mask, s_gpos[p], key, &w);
                    if (gp == 0xFFFFFFFFu) { full = true; w = false; }
                    if (w) won_mask |= 1u << e;
                    s_gpos[p] = gp & 0x7FFFFFFFu;
                }
            }
            __syncthreads();
            // phase 3: every use of a key learns its global entry
            if (emask)
                for (uint32_t m = emask; m; m &= m - 1u) {
                    const int e = __ffs((int) m) - 1;
                    const uint32_t ref = s_eref[e * 256 + threadIdx.x];
                    s_eref[e * 256 + threadIdx.x] = (ref & 0x80000000u) ? (ref & 0x7FFFFFFFu) : s_gpos[ref];
                }
        } else if (emask) {
            bx = vox[3 * (size_t) v]; by = vox[3 * (size_t) v + 1]; bz = vox[3 * (size_t) v + 2];
            // pass 1: start the table lines of all used edges on their way (the probes below are dependent chains)
            for (uint32_t m = emask; m; m &= m - 1u) {
                const int e = __ffs((int) m) - 1;
                float mx, my, mz;
                edge_midpoint(bx, by, bz, sx, sy, sz, e, mx, my, mz);
                uint32_t kx = __float_as_uint(mx);
                if (kx == 0xFFFFFFFFu) kx = 0x7FC00000u;
                const uint32_t pos0 = hash96(kx, __float_as_uint(my), __float_as_uint(mz)) & table_mask;
                prefetch_l2(table16 + pos0);
                s_eref[e * 256 + threadIdx.x] = pos0;
            }
            // pass 2: find-or-insert from the stored start position; remember the entry per edge and which ones this voxel created
            for (uint32_t m = emask; m; m &= m - 1u) {
                const int e = __ffs((int) m) - 1;
                float mx, my, mz;
                edge_midpoint(bx, by, bz, sx, sy, sz, e, mx, my, mz);
                uint32_t kx = __float_as_uint(mx);
                if (kx == 0xFFFFFFFFu) kx = 0x7FC00000u;   // keep the all-ones EMPTY pattern unreachable
                bool w;
                uint32_t pos = hash_probe_from(table16, table_mask, s_eref[e * 256 + threadIdx.x], kx, __float_as_uint(my), __float_as_uint(mz), 0xFFFFFFFEu, &w);
                if (pos == 0xFFFFFFFFu) { full = true; w = false; pos = 0; }
                if (w) won_mask |= 1u << e;
                s_eref[e * 256 + threadIdx.x] = pos;
            }
        }
        // ids of the vertices this tile created: block scan of the counts + one atomicAdd
        const uint32_t cnt = __popc(won_mask);
        const uint32_t incl = warp_inclusive_sum(cnt, lane);
        if (lane == 31u) s_w[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t run = 0;
            for (uint32_t w = 0; w < 8u; w++) { const uint32_t t = s_w[w]; s_w[w] = run; run += t; }
            s_w[8] = run;
            s_base = run ? atomicAdd(&st->n_uniq, run) : 0u;
        }
        __syncthreads();
        const uint32_t total = s_w[8], base = s_base;
        if (base + total > cap_uniq || base + total < base) {
            if (threadIdx.x == 0) atomicOr(&st->error_flags, ERR_UNIQ_CAP);
            continue;
        }
        if (emask) {
            const uint32_t rec = vparent ? vparent[v] : 0u;
            uint32_t uid = base + s_w[warp] + incl - cnt;
            for (uint32_t m = won_mask; m; m &= m - 1u) {
                const int e = __ffs((int) m) - 1;
                float mx, my, mz;
                edge_midpoint(bx, by, bz, sx, sy, sz, e, mx, my, mz);
                ustart[3 * (size_t) uid] = mx; ustart[3 * (size_t) uid + 1] = my; ustart[3 * (size_t) uid + 2] = mz;
                urec[uid] = rec;
                entry_uid[s_eref[e * 256 + threadIdx.x]] = uid;   // 4-byte side array of the table; readers come after the kernel boundary
                uid++;
            }
            const uint32_t ntri = mc.ntri[cube_index];
            const uint32_t t0 = tri_off[v];
            const unsigned long long packed = mc.packed[cube_index];
            for (uint32_t j = 0; j < 3 * ntri; j++) {
                const uint32_t e = (uint32_t) ((packed >> (4 * j)) & 0xFull);
                slot_ref[3 * (size_t) t0 + j] = s_eref[e * 256 + threadIdx.x];
            }
            for (uint32_t j = 0; j < ntri; j++) tri_rec[t0 + j] = rec;
        }
    }
    if (full) atomicOr(&st->error_flags, ERR_HASH_FULL);
    if (off_lattice) atomicOr(&st->error_flags, ERR_LATTICE);
}

// the mesh stage may be re-run on the same field: reset its counters and tickets (not error_flags, not the refine state)
__global__ void k_reset_mesh_state(DevState* st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    st->n_tris_raw = 0; st->n_uniq = 0; st->n_tris_out = 0; st->n_verts_out = 0;
    for (int i = TK_CLASSIFY; i < TK_COUNT; i++) st->ticket[i] = 0;
    st->n_stragglers = 0; st->weld_dups = 0;
    st->newton_iters = 0; st->tail_steps = 0; st->tail_rebuilds = 0; st->tail_unlisted = 0;
    st->n_escaped = 0; st->list_fallbacks = 0; st->n_orient_pending = 0;
    for (int i = WK_CLASSIFY; i < 6; i++) st->prim_evals[i] = 0;   // refine's counter is reset with the field
}

// Device-sized clears for the per-vertex / per-slot weld state (exactly as many entries as this mesh needs):
//   first_slot[0, n_uniq) = 0xFFFFFFFF, first_bits[0, ceil(3T/32)) = 0, weld table[0, weld_table_size(n_uniq)) = EMPTY.
__device__ __forceinline__ uint32_t weld_table_size(uint32_t n_uniq, uint32_t max_entries) {
    uint32_t s = 1024;
    while ((uint64_t) s * 4u < (uint64_t) n_uniq * 7u && s < max_entries) s <<= 1;   // load factor <= 4/7
    return s;
}
__global__ void __launch_bounds__(256) k_clear_weld_state(DevState* st, uint32_t* __restrict__ first_slot, uint32_t* __restrict__ first_bits,
                                                          uint4* __restrict__ table2, uint32_t max_entries, uint32_t cap_uniq, int clear_first_slot,
                                                          uint32_t* __restrict__ uesc) {
    const uint32_t nu = min(st->n_uniq, cap_uniq), T = st->n_tris_raw;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (clear_first_slot)
        for (uint32_t i = tid; i < nu; i += stride) first_slot[i] = 0xFFFFFFFFu;
    if (uesc)
        for (uint32_t i = tid; i < (nu + 31u) / 32u; i += stride) uesc[i] = 0u;
    const uint32_t nw = (3u * T + 31u) / 32u + 1u;
    for (uint32_t i = tid; i < nw; i += stride) first_bits[i] = 0u;
    const uint32_t ts = weld_table_size(nu, max_entries);
    const uint4 empty = make_uint4(SDM_HASH_EMPTY, SDM_HASH_EMPTY, SDM_HASH_EMPTY, SDM_HASH_EMPTY);
    for (uint32_t i = tid; i < ts; i += stride) table2[i] = empty;
    if (tid == 0) { st->n_tris_out = 0; st->n_verts_out = 0; st->ticket[TK_SCAN_FIRST] = 0; st->ticket[TK_SCAN_TRI] = 0; st->ticket[TK_PROJECT] = 0; st->n_stragglers = 0; st->ticket[TK_TAIL] = 0; st->weld_dups = 0; }
}

// closest_surface_point per distinct mid-point (signed_distance.cu:227-240), bulk phase: one lane per vertex; lanes that
// finish pull the next vertex (warp-level refill from a global ticket), so a warp's lanes stay busy although iteration
// counts differ.  A vertex that is still running after SDM_NEWTON_BULK_ITERS iterations is handed, with its state, to
// k_project_tail: the reference allows up to 10 000 iterations, and a lone slow lane would otherwise pin a whole warp
// (and drag its cell's primitives into the warp's mask) for thousands of latency-bound steps.
#define SDM_NEWTON_BULK_ITERS 40u
struct Straggler { uint32_t uid, it; float g[3]; float s[3]; uint32_t power, lam, stop_at, pad; };   // 48 B: vertex, iterate, Brent state

// 5 blocks of 128 threads per SM = 96 registers per thread: no spills, and 20 instead of 16 resident warps hide the low-ILP
// stretches (per-lane culling, Newton update) - measured 12 % faster than the unconstrained 128-register build; 6 blocks (80
// registers, 36 bytes spilled) is slower again: 3.15 -> 3.24 ms on configs[2]
#ifndef SDM_PROJ_MINB
#define SDM_PROJ_MINB 5
#endif
__global__ void __launch_bounds__(128, SDM_PROJ_MINB) k_project(const uint4* __restrict__ scene, DevState* st, const float* __restrict__ ustart,
                                                 float* __restrict__ upos, uint32_t cap_uniq, Straggler* __restrict__ stragglers,
                                                 uint32_t cap_stragglers, MaskGrid grid, uint32_t max_chunk,
                                                 const uint4* __restrict__ vl /* list records, or null: cell masks only */,
                                                 const uint32_t* __restrict__ urec, uint32_t* __restrict__ uesc, float slack2) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = min(st->n_uniq, cap_uniq);
    if (st->error_flags) return;
    const bool lists = vl != nullptr && sc.wmask != nullptr;
    bool have = false;
    bool inr = true;   // the lane's iterate is inside the region its vertex's list record is proven for
    uint32_t rec = 0, fallbacks = 0;
    uint32_t uid = 0, it = 0, iters_done = 0;
    unsigned long long work = 0;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    NewtonCycle cyc;
    cyc.start(0.f, 0.f, 0.f);
    // Vertices are taken in chunks of consecutive ids (consecutive ids are spatial neighbours, k_uid_offsets), one chunk
    // per warp at a time: the lanes of a warp then work in one neighbourhood, which keeps the tile's primitive list short.
    // chunk size: large enough for coherence, small enough that every warp gets >= ~4 chunks (load balance)
    const uint32_t warps_in_grid = (gridDim.x * blockDim.x) >> 5;
    const uint32_t CHUNK = sc.wmask ? min(max_chunk, max(32u, ((n / (warps_in_grid * 4u) + 31u) >> 5) << 5)) : 32u;   // coherence only matters when culling
    uint32_t chunk_next = 0, chunk_end = 0;
    bool drained = false;
    while (true) {
        // refill idle lanes: from the warp's current chunk, and from a fresh chunk in the same round when that one runs out
        for (int pass = 0; pass < 2; pass++) {
            const uint32_t need = __ballot_sync(0xffffffffu, !have);
            if (!need) break;
            if (chunk_next >= chunk_end) {
                if (drained) break;
                // guided hand-out: chunks shrink as the work runs out (half of an even share of what is left, at least one tile),
                // so that the warps finish together whatever the vertex count per warp is (shards of a multi-GPU run are small)
                uint32_t base = 0, len = 0;
                if (lane == 0) {
                    const uint32_t taken = min(*(volatile uint32_t*) &st->ticket[TK_PROJECT], n);
                    len = min(CHUNK, max(32u, (((n - taken) / (warps_in_grid * 2u)) + 31u) & ~31u));
                    base = atomicAdd(&st->ticket[TK_PROJECT], len);
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                len = __shfl_sync(0xffffffffu, len, 0);
                if (base >= n) { drained = true; break; }
                chunk_next = base; chunk_end = min(base + len, n);
#if SDM_PREFETCH
                // the chunk's start points and record indices on their way to L1 (32 entries = one 128-byte line of indices)
                if (base + lane * 32u < chunk_end) { if (lists) prefetch_l1(urec + base + lane * 32u); }
                if (3u * base + lane * 32u < 3u * chunk_end) prefetch_l1(ustart + 3 * (size_t) base + lane * 32u);
#endif
            }
            const uint32_t idx = chunk_next + __popc(need & ((1u << lane) - 1u));
            if (!have && idx < chunk_end) {
                uid = idx; it = 0; have = true;
                gx = ustart[3 * (size_t) uid]; gy = ustart[3 * (size_t) uid + 1]; gz = ustart[3 * (size_t) uid + 2];
                cyc.start(gx, gy, gz);
                inr = true;
                if (lists) rec = urec[uid];
            }
            chunk_next = min(chunk_next + (uint32_t) __popc(need), chunk_end);
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (drained || chunk_next >= chunk_end) { if (drained) break; }
            continue;   // fetch the next chunk
        }
        // The tile's primitive list: the union of the lanes' inherited records (valid while every iterate stays within `slack` of
        // its start point, which lies on the creating voxel: k_refine proved the records on the voxels inflated by that much),
        // else the cell masks at the lanes' current iterates.
        bool listed = false;
        if (lists && __all_sync(0xffffffffu, !have || inr)) listed = tile_list_from_records(sc, have, vl, rec, gx, gy, gz);
        if (!listed) { tile_mask_from_point(grid, sc, have, gx, gy, gz); fallbacks += lists ? 1u : 0u; }
        work += (unsigned long long) tile_prims(sc) * 13u * (uint32_t) __popc(__ballot_sync(0xffffffffu, have));
        if (have) {
            const bool collision = newton_step(sc, gx, gy, gz);
            it++;
            if (!collision) cyc.observe(gx, gy, gz, it);
            if (lists) {
                const float ex = gx - ustart[3 * (size_t) uid], ey = gy - ustart[3 * (size_t) uid + 1], ez = gz - ustart[3 * (size_t) uid + 2];
                inr = ex * ex + ey * ey + ez * ez <= slack2;   // NaN: outside
            }
            if (collision || it >= cyc.stop_at) {   // for (i = 0; !collision && i < 10000; i++)
                upos[3 * (size_t) uid] = gx; upos[3 * (size_t) uid + 1] = gy; upos[3 * (size_t) uid + 2] = gz;
                if (lists && !inr) { atomicOr(uesc + (uid >> 5), 1u << (uid & 31u)); atomicAdd(&st->n_escaped, 1u); }
                iters_done += it;
                have = false;
            } else if (it >= SDM_NEWTON_BULK_ITERS || !inr) {   // slow, or left its list's region: the tail kernel takes over (cell masks)
                const uint32_t slot = atomicAdd(&st->n_stragglers, 1u);
                if (slot < cap_stragglers) {
                    Straggler r;
                    r.uid = uid; r.it = it; r.g[0] = gx; r.g[1] = gy; r.g[2] = gz;
                    r.s[0] = cyc.sx; r.s[1] = cyc.sy; r.s[2] = cyc.sz; r.power = cyc.power; r.lam = cyc.lam;
                    r.stop_at = cyc.stop_at; r.pad = 0;
                    stragglers[slot] = r;
                    iters_done += it;
                    have = false;
                } else {
                    atomicSub(&st->n_stragglers, 1u);   // list full: keep iterating here
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) iters_done += __shfl_xor_sync(0xffffffffu, iters_done, o);
    if (lane == 0 && iters_done) atomicAdd(&st->newton_iters, (unsigned long long) iters_done);
    if (lane == 0 && work) atomicAdd(&st->prim_evals[WK_PROJECT], work);
    if (lane == 0 && fallbacks) atomicAdd(&st->list_fallbacks, fallbacks);
}

// Tail phase: one HALF-WARP per straggler.  The 13 evaluation points of a Newton step (the iterate and the 12 stencil
// points of empirical_normal) go to 13 lanes of the half-warp, so a step costs one evaluation's latency instead of
// thirteen; every lane then forms the same update from the 13 shuffled values (identical arithmetic => identical bits).
// Two vertices per warp keep 26 of 32 lanes busy when there are many stragglers (Mandelbulb: ~1 % of all vertices).
__global__ void __launch_bounds__(128) k_project_tail(const uint4* __restrict__ scene, DevState* st, float* __restrict__ upos,
                                                      const Straggler* __restrict__ stragglers, uint32_t cap_stragglers, MaskGrid grid,
                                                      const float* __restrict__ ustart, uint32_t* __restrict__ uesc /* null: no list records in use */, float slack2) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t hl = lane & 15u;            // lane within the half-warp
    const uint32_t half = lane >> 4;
    const uint32_t n = min(st->n_stragglers, cap_stragglers);
    unsigned long long extra_iters = 0, work = 0;
    uint32_t dbg_steps = 0, dbg_rebuilds = 0, dbg_unlisted = 0;
    while (true) {
        uint32_t idx0 = 0;
        if (lane == 0) idx0 = atomicAdd(&st->ticket[TK_TAIL], 2u);
        idx0 = __shfl_sync(0xffffffffu, idx0, 0);
        if (idx0 >= n) break;
        const uint32_t idx = idx0 + half;
        bool running = idx < n;
        Straggler r;
        r.uid = 0; r.it = 0; r.g[0] = r.g[1] = r.g[2] = 0.f; r.s[0] = r.s[1] = r.s[2] = 0.f; r.power = 1; r.lam = 0; r.stop_at = 0; r.pad = 0;
        if (running) r = stragglers[idx];
        float gx = r.g[0], gy = r.g[1], gz = r.g[2];
        uint32_t it = r.it;
        NewtonCycle cyc;
        cyc.sx = r.s[0]; cyc.sy = r.s[1]; cyc.sz = r.s[2]; cyc.power = r.power; cyc.lam = r.lam; cyc.stop_at = r.stop_at;
        running = running && it < cyc.stop_at;
        // The tile's primitive list is built for a ball of extra radius `slack` around each iterate and kept until an iterate
        // leaves its ball (or a half finishes): a slow orbit moves ~1e-4 per step, so the list is rebuilt every ~100 steps
        // instead of every step.  (A larger ball only makes the list a superset: still exact.)  An orbit that keeps JUMPING -
        // between the two sides of a crease, say - would rebuild on every step (an animated frame spent 157 of its 171 ms on one such
        // vertex): when rebuilds come in quick succession the ball doubles, up to what the cell look-up covers, until it holds the orbit.
        float slack = 0.01f;
        const float max_slack = grid.enabled ? 1.5f * grid.cell : 0.01f;
        uint32_t recent = 0, since = 0;
        float lcx = gx, lcy = gy, lcz = gz;
        bool list_valid = false;
        uint32_t listed_for = 0;   // the halves the current list was built for
        while (true) {
            const uint32_t run_mask = __ballot_sync(0xffffffffu, running);
            if (!run_mask) break;
            // a half that has finished must not leave its primitives in the other's list (a vertex that went NaN outside the mask
            // grid left the whole table there: 1 400 steps of its partner at 106 us each - the 171 ms frame of the animated run)
            if (run_mask != listed_for) list_valid = false;
            const float mvx = gx - lcx, mvy = gy - lcy, mvz = gz - lcz;
            const bool moved = running && !(mvx * mvx + mvy * mvy + mvz * mvz <= (0.5f * slack) * (0.5f * slack));   // NaN -> rebuild
            if (!list_valid || __any_sync(0xffffffffu, moved)) {
                if (list_valid && ++recent >= 4u && slack < max_slack) { slack = fminf(2.0f * slack, max_slack); recent = 0; }
                listed_for = run_mask;
#ifdef SDM_TAIL_SERIAL_REFINE
                tile_mask_from_point(grid, sc, running && hl == 0, gx, gy, gz, slack);
#else
                tile_mask_from_half_points(grid, sc, running, gx, gy, gz, slack);
#endif
                lcx = gx; lcy = gy; lcz = gz;
                list_valid = true;
                dbg_rebuilds++;
            }
            dbg_steps++;
            if (sc.wmask && *sc.tcount == SDM_TLIST_NONE) dbg_unlisted++;
            if (++since >= 64u) { since = 0; recent >>= 1; }   // rebuilds far apart do not add up
            if (lane == 0) work += (unsigned long long) tile_prims(sc) * 13u * (uint32_t) __popc(__ballot_sync(0xffffffffu, running && hl == 0));
            else (void) __ballot_sync(0xffffffffu, running && hl == 0);
            // half-lane 0: the iterate; half-lane 1 + 4*axis + s: stencil point s of that axis (same construction as normal_points)
            float x = gx, y = gy, z = gz;
            if (hl >= 1 && hl <= 12) {
                const uint32_t q = hl - 1u, a = q >> 2, sidx = q & 3u;
                const float o = sidx == 0 ? 2.0f * SDM_NORMAL_EPSILON : (sidx == 1 ? SDM_NORMAL_EPSILON : (sidx == 2 ? -SDM_NORMAL_EPSILON : -2.0f * SDM_NORMAL_EPSILON));
                x = gx + (a == 0 ? o : 0.0f); y = gy + (a == 1 ? o : 0.0f); z = gz + (a == 2 ? o : 0.0f);
            }
            float fv = 0.0f;
            if (running && hl <= 12) fv = eval_scene1(sc, x, y, z);
            float f[13];
#pragma unroll
            for (int j = 0; j < 13; j++) f[j] = __shfl_sync(0xffffffffu, fv, j, 16);
            if (running) {
                float nx, ny, nz;
                normal_from_samples<1>(f, nx, ny, nz);
                const float sd = f[0];
                gx -= sd * nx; gy -= sd * ny; gz -= sd * nz;
                const bool collision = fabsf(sd) <= 0.00001f;
                it++;
                if (!collision) cyc.observe(gx, gy, gz, it);
                running = !collision && it < cyc.stop_at;
            }
        }
        if (idx < n && hl == 0) {
            upos[3 * (size_t) r.uid] = gx; upos[3 * (size_t) r.uid + 1] = gy; upos[3 * (size_t) r.uid + 2] = gz;
            extra_iters += it - r.it;
            if (uesc) {   // the later per-vertex / per-triangle kernels must not use this vertex's list record if it ended outside its region
                const float ex = gx - ustart[3 * (size_t) r.uid], ey = gy - ustart[3 * (size_t) r.uid + 1], ez = gz - ustart[3 * (size_t) r.uid + 2];
                if (!(ex * ex + ey * ey + ez * ez <= slack2)) { atomicOr(uesc + (r.uid >> 5), 1u << (r.uid & 31u)); atomicAdd(&st->n_escaped, 1u); }
            }
        }
    }
    extra_iters += __shfl_xor_sync(0xffffffffu, extra_iters, 16);
    if (lane == 0 && extra_iters) atomicAdd(&st->newton_iters, extra_iters);
    if (lane == 0 && work) atomicAdd(&st->prim_evals[WK_TAIL], work);
    if (lane == 0 && dbg_steps) {
        atomicAdd(&st->tail_steps, (unsigned long long) dbg_steps); atomicAdd(&st->tail_rebuilds, (unsigned long long) dbg_rebuilds);
        atomicAdd(&st->tail_unlisted, (unsigned long long) dbg_unlisted);
    }
}

__device__ __forceinline__ uint32_t tile_group(uint32_t ntiles, uint32_t warps) { return max(1u, min(8u, ntiles / (warps * 4u))); }
// Inserts vertex u's quantised weld key (src/cuda/mod.rs:270) into the key table with value 0xFFFFFFFF ("no slot yet");
// returns the entry, counts vertices that met an existing key.
__device__ __forceinline__ uint32_t weld_insert_key(DevState* st, const float* __restrict__ upos, uint32_t u, uint4* table, uint32_t table_mask) {
    const uint32_t kx = weld_key_component(upos[3 * (size_t) u]);
    const uint32_t ky = weld_key_component(upos[3 * (size_t) u + 1]);
    const uint32_t kz = weld_key_component(upos[3 * (size_t) u + 2]);
    bool won;
    const uint32_t pos = hash_find_or_insert(table, table_mask, kx, ky, kz, 0xFFFFFFFFu, &won);
    if (pos == 0xFFFFFFFFu) atomicOr(&st->error_flags, ERR_HASH_FULL);
    else if (!won) atomicAdd(&st->weld_dups, 1u);
    return pos;
}

// empirical_normal per projected vertex.  weld_table != nullptr: the vertex's weld key is inserted here as well - a random
// DRAM access per vertex that is free while the SM is busy with the twelve evaluations (as a kernel of its own it took a
// third of this kernel's time doing nothing but waiting for memory).
#ifndef SDM_NRM_MINB
#define SDM_NRM_MINB 6   /* 80 registers, no spills since the per-lane test takes one candidate per step: 1.24 -> 1.16 ms on configs[2] (96 registers / 5 blocks was the better choice while that test was unrolled by two) */
#endif
__global__ void __launch_bounds__(128, SDM_NRM_MINB) k_vertex_normals(const uint4* __restrict__ scene, DevState* st, const float* __restrict__ upos,
                                                        float* __restrict__ unrm, uint32_t cap_uniq, MaskGrid grid, uint4* weld_table,
                                                        uint32_t weld_max_entries, uint32_t* __restrict__ wref,
                                                        const uint4* __restrict__ vl, const uint32_t* __restrict__ urec, const uint32_t* __restrict__ uesc) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t n = min(st->n_uniq, cap_uniq);
    if (st->error_flags) return;
    const bool lists = vl != nullptr && sc.wmask != nullptr;
    uint32_t fallbacks = 0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t table_mask = weld_table_size(n, weld_max_entries) - 1u;   // same size k_clear_weld_state cleared
    unsigned long long work = 0;
    // tiles are handed out dynamically, a few at a time (every warp should still get >= ~4 hand-outs): their cost varies with the list
    // lengths, and a static split left a tail
    uint32_t group = tile_group((n + 31u) >> 5, warps_total);
    for (uint32_t g0 = 0, gk = group;; gk++) {
        if (gk == group) {
            if (lane == 0) {   // guided: the groups shrink as the work runs out
                const uint32_t taken = min(*(volatile uint32_t*) &st->ticket[TK_NORMALS], n);
                group = max(1u, min(8u, (n - taken) / (warps_total * 64u)));
                g0 = atomicAdd(&st->ticket[TK_NORMALS], 32u * group);
            }
            g0 = __shfl_sync(0xffffffffu, g0, 0);
            group = __shfl_sync(0xffffffffu, group, 0);
            gk = 0;
        }
        const uint32_t u0 = g0 + 32u * gk;
        if (u0 >= n) { if (gk == 0) break; gk = group - 1; continue; }
        const uint32_t u = u0 + lane;
        const bool active = u < n;
#if SDM_PREFETCH
        if (gk + 1 < group && u0 + 32u < n) {   // the next tile of this group
            if (lane < 3u) prefetch_l1(upos + 3 * (size_t) (u0 + 32u) + lane * 32u);
            else if (lane == 3u && lists) prefetch_l1(urec + u0 + 32u);
        }
#endif
        float x = 0.f, y = 0.f, z = 0.f;
        if (active) { x = upos[3 * (size_t) u]; y = upos[3 * (size_t) u + 1]; z = upos[3 * (size_t) u + 2]; }
        if (weld_table && active) wref[u] = weld_insert_key(st, upos, u, weld_table, table_mask);
        bool listed = false;
        if (lists) {
            const bool esc = active && ((uesc[u >> 5] >> (u & 31u)) & 1u);
            if (!__any_sync(0xffffffffu, esc)) listed = tile_list_from_records(sc, active, vl, active ? urec[u] : 0u, x, y, z);
        }
        if (!listed) { tile_mask_from_point(grid, sc, active, x, y, z); fallbacks += lists ? 1u : 0u; }
        work += (unsigned long long) tile_prims(sc) * 12u * min(32u, n - u0);
        if (active) {
            float nx, ny, nz;
            empirical_normal(sc, x, y, z, nx, ny, nz);
            unrm[3 * (size_t) u] = nx; unrm[3 * (size_t) u + 1] = ny; unrm[3 * (size_t) u + 2] = nz;
        }
    }
    if (lane == 0 && work) atomicAdd(&st->prim_evals[WK_NORMALS], work);
    if (lane == 0 && fallbacks) atomicAdd(&st->list_fallbacks, fallbacks);
}

// Per raw triangle: orientation test and the reference host's triangle filter.
// One triangle's loads: vertex ids (pre-flip), positions, centroid (compute_mesh_generation.cu:104).
struct OrientTri { uint32_t u[3]; float v[3][3]; float mx, my, mz; };
__device__ __forceinline__ void orient_load(OrientTri& r, bool active, uint32_t t, const uint32_t* __restrict__ entry_uid, const uint32_t* __restrict__ slot_ref,
                                            const float* __restrict__ upos) {
#pragma unroll
    for (int j = 0; j < 3; j++) { r.u[j] = 0; r.v[j][0] = r.v[j][1] = r.v[j][2] = 0.f; }
    r.mx = r.my = r.mz = 0.f;
    if (!active) return;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        r.u[j] = entry_uid[slot_ref[3 * (size_t) t + j]];
        r.v[j][0] = upos[3 * (size_t) r.u[j]]; r.v[j][1] = upos[3 * (size_t) r.u[j] + 1]; r.v[j][2] = upos[3 * (size_t) r.u[j] + 2];
    }
    // (v0 + v1 + v2) / 3.0f
    r.mx = (r.v[0][0] + r.v[1][0] + r.v[2][0]) / 3.0f; r.my = (r.v[0][1] + r.v[1][1] + r.v[2][1]) / 3.0f; r.mz = (r.v[0][2] + r.v[1][2] + r.v[2][2]) / 3.0f;
}
// normalize(cross(v1 - v0, v2 - v0))   (compute_mesh_generation.cu:103)
__device__ __forceinline__ void orient_face_normal(const OrientTri& r, float& tnx, float& tny, float& tnz) {
    const float ax = r.v[1][0] - r.v[0][0], ay = r.v[1][1] - r.v[0][1], az = r.v[1][2] - r.v[0][2];
    const float bx = r.v[2][0] - r.v[0][0], by = r.v[2][1] - r.v[0][1], bz = r.v[2][2] - r.v[0][2];
    const float cx = ay * bz - by * az, cy = az * bx - bz * ax, cz = ax * by - bx * ay;
    const float inv = 1.0f / sqrtf(dot3(cx, cy, cz, cx, cy, cz));
    tnx = cx * inv; tny = cy * inv; tnz = cz * inv;
}
// the triangle's corner order after the flip (:105-113), the reference host's finite filter (src/cuda/mod.rs:289), first-occurrence slots
__device__ __forceinline__ bool orient_store(const OrientTri& r, bool flip, uint32_t t, uint32_t* __restrict__ tri_uid, uint32_t* __restrict__ first_slot) {
    const uint32_t f0 = flip ? r.u[2] : r.u[0], f2 = flip ? r.u[0] : r.u[2];
    const float first_x = flip ? r.v[2][0] : r.v[0][0];
    tri_uid[3 * (size_t) t] = f0; tri_uid[3 * (size_t) t + 1] = r.u[1]; tri_uid[3 * (size_t) t + 2] = f2;
    const bool valid = fabsf(first_x) <= FLT_MAX;   // kept iff vertices[0].position.x is finite
    if (valid) {
        atomicMin(first_slot + f0, 3u * t);
        atomicMin(first_slot + r.u[1], 3u * t + 1u);
        atomicMin(first_slot + f2, 3u * t + 2u);
    }
    return valid;
}
// The tile's primitive list for the centroids.  The centroid lies in the convex hull of the three vertices; each ended within `slack`
// of its start point on this triangle's voxel (unless flagged), so the centroid is inside the region the voxel's list record is proven for.
__device__ __forceinline__ bool orient_tile_list(const SceneView& sc, const MaskGrid& grid, bool lists, bool active, const OrientTri& r, uint32_t t,
                                                 const uint4* __restrict__ vl, const uint32_t* __restrict__ tri_rec, const uint32_t* __restrict__ uesc, bool& esc) {
    bool listed = false;
    esc = false;
    if (lists) {
        if (active) esc = (((uesc[r.u[0] >> 5] >> (r.u[0] & 31u)) | (uesc[r.u[1] >> 5] >> (r.u[1] & 31u)) | (uesc[r.u[2] >> 5] >> (r.u[2] & 31u))) & 1u) != 0u;
        if (!__any_sync(0xffffffffu, esc)) listed = tile_list_from_records(sc, active, vl, active ? tri_rec[t] : 0u, r.mx, r.my, r.mz);
    }
    if (!listed) tile_mask_from_point(grid, sc, active, r.mx, r.my, r.mz);
    return listed;
}

// QUICK = false: the reference's statement - twelve evaluations per triangle.
// QUICK = true (scenes of 1-Lipschitz primitives: everything but the Mandelbulb estimator): only the SIGN of dot(tn, n) is used (:105),
// n = normalize(d), d_axis = 8 A_axis - B_axis with A = f(+e) - f(-e), B = f(+2e) - f(-2e) (signed_distance.cu:186-199).  The exact
// scene function is 1-Lipschitz, so |B_axis| <= 4e (+ what rounding adds) whatever the scene does between the samples, and
//     8 |tn . A|  >  |tn|_1 * Bmax   ==>   sign(tn . d) = sign(tn . A):
// six evaluations decide the triangle.  What rounding adds, all in the bound: the evaluated f differs from the exact fold by at most
// (L + 6) * 2e-7 * V (L = folded primitives, V = largest distance magnitude <= |centroid|_1 + `reach`, the host's bound on
// |centre| + extent over the table: a distance rounds to <= 6 units of 2e-7 V, and a smooth-min step is non-expansive in the max norm
// and rounds its own operations to one unit); the sample points are 4e apart up to 1e-3 relative; the reference's own left-to-right
// sum rounds to 4e-6 (max|f| + 1e-3); 7.9 instead of 8 covers this test's own arithmetic.  The margin left makes |cos(tn, n)| > 1e-3,
// far above the rounding of the reference's normalize and dot.  Triangles the test leaves open (a NaN anywhere, a sliver whose normal
// is more than ~60 degrees off the gradient, a vertex flagged as escaped) go to `pending`: k_orient_pending applies the full statement.
#ifndef SDM_ORIENT_MINB
#define SDM_ORIENT_MINB 6   /* 80 registers (20 bytes spilled): 1.385 -> 1.315 ms on configs[2]; the six-sample pass needs fewer registers than k_project's 13 points */
#endif
template <bool QUICK>
__global__ void __launch_bounds__(128, SDM_ORIENT_MINB) k_orient(const uint4* __restrict__ scene, DevState* st, const uint32_t* __restrict__ entry_uid,
                                                const uint32_t* __restrict__ slot_ref, const float* __restrict__ upos,
                                                uint32_t* __restrict__ tri_uid, uint32_t* __restrict__ first_slot,
                                                uint32_t* __restrict__ tri_valid_bits, MaskGrid grid,
                                                const uint4* __restrict__ vl, const uint32_t* __restrict__ tri_rec, const uint32_t* __restrict__ uesc,
                                                uint32_t* __restrict__ pending, float reach) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t T = st->n_tris_raw;
    if (st->error_flags) return;
    const bool lists = vl != nullptr && sc.wmask != nullptr;
    uint32_t fallbacks = 0;
    const uint32_t lane = threadIdx.x & 31u;
    // warp-contiguous mapping so that one lane can write the 32 validity bits of a warp's triangles
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    unsigned long long work = 0;
    uint32_t group = tile_group((T + 31u) >> 5, warps_total);
    for (uint32_t g0 = 0, gk = group;; gk++) {   // dynamic, guided hand-out, see k_vertex_normals
        if (gk == group) {
            if (lane == 0) {
                const uint32_t taken = min(*(volatile uint32_t*) &st->ticket[TK_ORIENT], T);
                group = max(1u, min(8u, (T - taken) / (warps_total * 64u)));
                g0 = atomicAdd(&st->ticket[TK_ORIENT], 32u * group);
            }
            g0 = __shfl_sync(0xffffffffu, g0, 0);
            group = __shfl_sync(0xffffffffu, group, 0);
            gk = 0;
        }
        const uint32_t t0 = g0 + 32u * gk;
        if (t0 >= T) { if (gk == 0) break; gk = group - 1; continue; }
        const uint32_t t = t0 + lane;
#if SDM_PREFETCH
        if (gk + 1 < group && t0 + 32u < T) {   // the next tile of this group: its table references and records
            if (lane < 3u) prefetch_l1(slot_ref + 3 * (size_t) (t0 + 32u) + lane * 32u);
            else if (lane == 3u && lists) prefetch_l1(tri_rec + t0 + 32u);
        }
#endif
        OrientTri r;
        orient_load(r, t < T, t, entry_uid, slot_ref, upos);
        bool esc;
        const bool listed = orient_tile_list(sc, grid, lists, t < T, r, t, vl, tri_rec, uesc, esc);
        fallbacks += (lists && !listed) ? 1u : 0u;
        const uint32_t L = tile_prims(sc);
        work += (unsigned long long) L * (QUICK ? 6u : 12u) * min(32u, T - t0);
        bool valid = false, open = false;
        if (t < T) {
            float tnx, tny, tnz;
            orient_face_normal(r, tnx, tny, tnz);
            if (QUICK) {
                // the +e / -e samples of each axis, built as normal_points builds them
                float px[6], py[6], pz[6], f[6];
#pragma unroll
                for (int a = 0; a < 3; a++)
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const float o = q == 0 ? SDM_NORMAL_EPSILON : -SDM_NORMAL_EPSILON;
                        px[2 * a + q] = r.mx + (a == 0 ? o : 0.0f); py[2 * a + q] = r.my + (a == 1 ? o : 0.0f); pz[2 * a + q] = r.mz + (a == 2 ? o : 0.0f);
                    }
                eval_scene<6>(sc, px, py, pz, f);
                const float dotA = tnx * (f[0] - f[1]) + tny * (f[2] - f[3]) + tnz * (f[4] - f[5]);
                const float l1 = fabsf(tnx) + fabsf(tny) + fabsf(tnz);
                const float fmax = fmaxf(fmaxf(fmaxf(fabsf(f[0]), fabsf(f[1])), fmaxf(fabsf(f[2]), fabsf(f[3]))), fmaxf(fabsf(f[4]), fabsf(f[5])));
                const float unit = 2e-7f * (reach + fabsf(r.mx) + fabsf(r.my) + fabsf(r.mz));
                const float bmax = 4.0f * SDM_NORMAL_EPSILON * 1.003f + 2.0f * (float) (L + 6u) * unit + 4e-6f * (fmax + 1e-3f);
                if (!esc && 7.9f * fabsf(dotA) > l1 * bmax) valid = orient_store(r, dotA < 0.0f, t, tri_uid, first_slot);   // false for any NaN / inf
                else open = true;
            } else {
                float nx, ny, nz;
                empirical_normal(sc, r.mx, r.my, r.mz, nx, ny, nz);   // empirical_normal(sd_obj, centroid)   (:104)
                valid = orient_store(r, dot3(tnx, tny, tnz, nx, ny, nz) <= 0.0f, t, tri_uid, first_slot);   // :105
            }
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) tri_valid_bits[t0 >> 5] = bits;
        if (QUICK) {
            const uint32_t ob = __ballot_sync(0xffffffffu, open);
            if (ob) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&st->n_orient_pending, (uint32_t) __popc(ob));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (open) pending[base + __popc(ob & ((1u << lane) - 1u))] = t;   // capacity: one slot per raw triangle
            }
        }
    }
    if (lane == 0 && work) atomicAdd(&st->prim_evals[WK_ORIENT], work);
    if (lane == 0 && fallbacks) atomicAdd(&st->list_fallbacks, fallbacks);
}

// The triangles k_orient<true> left open: the reference's full statement.  Their validity bits are OR-ed into the words k_orient wrote.
__global__ void __launch_bounds__(128) k_orient_pending(const uint4* __restrict__ scene, DevState* st, const uint32_t* __restrict__ entry_uid,
                                                        const uint32_t* __restrict__ slot_ref, const float* __restrict__ upos,
                                                        uint32_t* __restrict__ tri_uid, uint32_t* __restrict__ first_slot,
                                                        uint32_t* __restrict__ tri_valid_bits, MaskGrid grid,
                                                        const uint4* __restrict__ vl, const uint32_t* __restrict__ tri_rec, const uint32_t* __restrict__ uesc,
                                                        const uint32_t* __restrict__ pending) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t n = min(st->n_orient_pending, st->n_tris_raw);
    if (st->error_flags) return;
    const bool lists = vl != nullptr && sc.wmask != nullptr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long work = 0;
    for (uint32_t i0 = warp_id * 32u; i0 < n; i0 += warps_total * 32u) {
        const bool active = i0 + lane < n;
        const uint32_t t = active ? pending[i0 + lane] : 0u;
        OrientTri r;
        orient_load(r, active, t, entry_uid, slot_ref, upos);
        bool esc;
        orient_tile_list(sc, grid, lists, active, r, t, vl, tri_rec, uesc, esc);
        work += (unsigned long long) tile_prims(sc) * 12u * min(32u, n - i0);
        if (active) {
            float tnx, tny, tnz, nx, ny, nz;
            orient_face_normal(r, tnx, tny, tnz);
            empirical_normal(sc, r.mx, r.my, r.mz, nx, ny, nz);
            if (orient_store(r, dot3(tnx, tny, tnz, nx, ny, nz) <= 0.0f, t, tri_uid, first_slot)) atomicOr(tri_valid_bits + (t >> 5), 1u << (t & 31u));
        }
    }
    if (lane == 0 && work) atomicAdd(&st->prim_evals[WK_ORIENT], work);
}

// The reference-order weld (src/cuda/mod.rs:263-296).  A vertex's key entry is created by weld_insert_key (fused into
// k_vertex_normals, or k_weld_keys for a merged multi-shard list); first_slot[u] = smallest slot id (3*triangle + corner,
// post-flip order) at which vertex u occurs in a kept triangle (k_orient).
//   * weld_dups == 0 (the usual case): all keys are distinct, every referenced vertex is its key's first occurrence, and no
//     kernel below touches the key table again;
//   * otherwise k_weld_min leaves each key's smallest slot in its entry, and the vertex whose first slot equals that
//     minimum is the key's first occurrence in the reference's scan order.
__global__ void __launch_bounds__(256) k_weld_keys(DevState* st, const float* __restrict__ upos, uint4* table, uint32_t max_entries,
                                                   uint32_t* __restrict__ wref, uint32_t cap_uniq) {
    const uint32_t n = min(st->n_uniq, cap_uniq);
    if (st->error_flags) return;
    const uint32_t table_mask = weld_table_size(n, max_entries) - 1u;   // same size k_clear_weld_state cleared
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) wref[u] = weld_insert_key(st, upos, u, table, table_mask);
}
__global__ void __launch_bounds__(256) k_weld_min(DevState* st, const uint32_t* __restrict__ first_slot, uint4* table,
                                                  const uint32_t* __restrict__ wref, uint32_t cap_uniq) {
    const uint32_t n = min(st->n_uniq, cap_uniq);
    if (st->error_flags || st->weld_dups == 0u) return;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) {
        const uint32_t fs = first_slot[u];
        if (fs != 0xFFFFFFFFu) atomicMin(reinterpret_cast<uint32_t*>(table + wref[u]) + 3, fs);
    }
}
// marks the slot of every key's first occurrence
__global__ void __launch_bounds__(256) k_weld_mark(DevState* st, const uint32_t* __restrict__ first_slot, const uint4* __restrict__ table,
                                                   const uint32_t* __restrict__ wref, uint32_t* first_bits, uint32_t cap_uniq) {
    const uint32_t n = min(st->n_uniq, cap_uniq);
    if (st->error_flags) return;
    const bool dups = st->weld_dups != 0u;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) {
        const uint32_t fs = first_slot[u];
        if (fs == 0xFFFFFFFFu) continue;
        if (!dups || reinterpret_cast<const uint32_t*>(table + wref[u])[3] == fs) atomicOr(first_bits + (fs >> 5), 1u << (fs & 31u));
    }
}

// Exclusive prefix pop-count over a bit mask: word_prefix[w] = number of set bits in words [0, w).
// which: 0 -> bits cover 3*n_tris_raw slots, total to n_verts_out; 1 -> bits cover n_tris_raw, total to n_tris_out.
// which: 2 / 3 -> the peer exchange's bitmaps: number of bits read from *nbits_ptr (+ 64 spare bits, so that a rank look-up one
//        word past the end is defined), total to st->peer_scan_total[which - 2]; the ticket must have been reset (k_peer_scan_reset).
__global__ void __launch_bounds__(256) k_bitscan(DevState* st, const uint32_t* __restrict__ bits, uint32_t* __restrict__ word_prefix,
                                                 int which, uint32_t epoch, uint64_t* tiles, const uint32_t* nbits_ptr = nullptr) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_w[10];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t nbits = which >= 2 ? *nbits_ptr + 64u : (which == 0 ? 3u * st->n_tris_raw : st->n_tris_raw);
    const uint32_t nwords = (nbits + 31u) >> 5;
    const uint32_t per_tile = blockDim.x * 4u;   // 4 consecutive words per thread
    const uint32_t ntiles = (nwords + per_tile - 1u) / per_tile;
    uint32_t* total_out = which >= 2 ? &st->peer_scan_total[which - 2] : (which == 0 ? &st->n_verts_out : &st->n_tris_out);
    const int tk = (which & 1) == 0 ? TK_SCAN_FIRST : TK_SCAN_TRI;
    const bool bad = which < 2 && st->error_flags != 0;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&st->ticket[tk], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles || bad) {
            if ((tile == 0 || bad) && threadIdx.x == 0) *total_out = 0;
            break;
        }
        const uint32_t w0 = tile * per_tile + threadIdx.x * 4u;
        uint32_t c[4] = { 0, 0, 0, 0 };
        if (w0 + 3u < nwords) {   // (the bit arrays are padded by 32 words, but stay within the words that were cleared)
            const uint4 q = *reinterpret_cast<const uint4*>(bits + w0);
            c[0] = __popc(q.x); c[1] = __popc(q.y); c[2] = __popc(q.z); c[3] = __popc(q.w);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (w0 + j < nwords) c[j] = __popc(bits[w0 + j]);
        }
        const uint32_t mine = c[0] + c[1] + c[2] + c[3];
        const uint32_t incl = warp_inclusive_sum(mine, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t end;
        const uint32_t base = block_lookback(tiles, tile, epoch, total, s_w, end);
        uint32_t o = base + incl - mine;
        if (w0 + 3u < nwords) {
            *reinterpret_cast<uint4*>(word_prefix + w0) = make_uint4(o, o + c[0], o + c[0] + c[1], o + c[0] + c[1] + c[2]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) { if (w0 + j < nwords) word_prefix[w0 + j] = o; o += c[j]; }
        }
        if (tile == ntiles - 1 && threadIdx.x == 0) *total_out = end;
    }
}

__device__ __forceinline__ uint32_t bit_rank(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ word_prefix, uint32_t i) {
    return word_prefix[i >> 5] + (uint32_t) __popc(bits[i >> 5] & ((1u << (i & 31u)) - 1u));
}

// vidx[u] = output index of vertex u (rank of its key's first slot); the first occurrence also writes position AND normal
// (src/cuda/mod.rs:279-284)
__global__ void __launch_bounds__(256) k_emit_vertices(DevState* st, const uint32_t* __restrict__ first_slot, const uint4* __restrict__ table,
                                                       const uint32_t* __restrict__ wref, const uint32_t* __restrict__ first_bits,
                                                       const uint32_t* __restrict__ first_prefix, const float* __restrict__ upos,
                                                       const float* __restrict__ unrm, float* __restrict__ out_pos, float* __restrict__ out_nrm,
                                                       uint32_t* __restrict__ vidx, uint32_t cap_uniq) {
    const uint32_t n = min(st->n_uniq, cap_uniq);
    if (st->error_flags) return;
    const bool dups = st->weld_dups != 0u;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) {
        const uint32_t fs = first_slot[u];
        if (fs == 0xFFFFFFFFu) continue;   // not referenced by a kept triangle: never read through vidx either
        const uint32_t kfs = dups ? reinterpret_cast<const uint32_t*>(table + wref[u])[3] : fs;
        const uint32_t rank = bit_rank(first_bits, first_prefix, kfs);
        vidx[u] = rank;
        if (kfs != fs) continue;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            out_pos[3 * (size_t) rank + c] = upos[3 * (size_t) u + c];
            out_nrm[3 * (size_t) rank + c] = unrm[3 * (size_t) u + c];
        }
    }
}
__global__ void __launch_bounds__(256) k_emit_indices(DevState* st, const uint32_t* __restrict__ tri_uid, const uint32_t* __restrict__ vidx,
                                                      const uint32_t* __restrict__ tri_valid_bits, const uint32_t* __restrict__ tri_prefix,
                                                      uint32_t* __restrict__ out_idx) {
    const uint32_t T = st->n_tris_raw;
    if (st->error_flags) return;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
        if (!((tri_valid_bits[t >> 5] >> (t & 31u)) & 1u)) continue;
        const uint32_t ot = bit_rank(tri_valid_bits, tri_prefix, t);
#pragma unroll
        for (int j = 0; j < 3; j++) out_idx[3 * (size_t) ot + j] = vidx[tri_uid[3 * (size_t) t + j]];
    }
}

// The reference's device output format: 5 Triangle slots per voxel (compute_mesh_generation.cu:71-72), post-flip
// vertices (:111-113); unused slots are `{ POINT_NAN, POINT_NAN }` = vertex 0 NaN, vertices 1..2 zero (:116-118).
__global__ void __launch_bounds__(256) k_soup(DevState* st, int level, const uint8_t* __restrict__ cases, const uint32_t* __restrict__ tri_off,
                                              const uint32_t* __restrict__ tri_uid, const float* __restrict__ upos, const float* __restrict__ unrm,
                                              float* __restrict__ out /* n*5*18 */) {
    const uint32_t n = st->level_count[level];
    if (st->error_flags) return;
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        const uint32_t ntri = c_mc_ntri[cases[v]];
        const uint32_t t0 = tri_off[v];
        float* o = out + (size_t) v * 90;
        for (uint32_t t = 0; t < 5; t++) {
            if (t < ntri) {
                for (int j = 0; j < 3; j++) {
                    const uint32_t u = tri_uid[3 * (size_t) (t0 + t) + j];
                    for (int c = 0; c < 3; c++) {
                        o[t * 18 + j * 6 + c] = upos[3 * (size_t) u + c];
                        o[t * 18 + j * 6 + 3 + c] = unrm[3 * (size_t) u + c];
                    }
                }
            } else {
                for (int q = 0; q < 18; q++) o[t * 18 + q] = q < 6 ? __int_as_float(0x7fc00000) : 0.0f;
            }
        }
    }
}

// ---- shards (multi-GPU) ----------------------------------------------------------------------------------
// Restrict the list of `level` to the contiguous part [n*shard/count, n*(shard+1)/count) (64-bit arithmetic),
// copied to the front of the other ping-pong buffer.  Children keep their parent's order (compute_mesh_generation.cu:51)
// and level 0 is x-major (src/cuda/mod.rs:110-119), so a contiguous part of the list is an x-slab at every level.
// Shard bounds in the DENSE level-0 list by the cells' weights (k_build_masks: 0 = provably no surface, else the number of primitives
// the cell keeps): every shard gets the same total weight, so that no level has to be refined redundantly to find a balanced split
// and shards in dense regions - longer primitive lists per evaluation - get fewer voxels.  bounds[0..1] = [lo, hi); the same
// arithmetic on every rank, so the shards tile the list.
__global__ void __launch_bounds__(1024) k_shard_bounds_by_flags(const uint8_t* __restrict__ flags, uint32_t n, uint32_t shard, uint32_t count, uint32_t* bounds,
                                                                const float* __restrict__ frac /* cumulative weight fractions [count + 1], or null: equal shares */) {
    __shared__ unsigned long long s_part[1024];
    const uint32_t per = (n + 1023u) / 1024u, i0 = min(threadIdx.x * per, n), i1 = min(i0 + per, n);
    unsigned long long sum = 0;
    if ((per & 15u) == 0u && i1 - i0 == per && (reinterpret_cast<uintptr_t>(flags) & 15u) == 0u) {   // 16 weights per load
        const uint4* f4 = reinterpret_cast<const uint4*>(flags + i0);
        for (uint32_t q = 0; q < per / 16u; q++) {
            const uint4 w = __ldg(f4 + q);
            sum += __dp4a(w.x, 0x01010101u, 0u) + __dp4a(w.y, 0x01010101u, 0u) + __dp4a(w.z, 0x01010101u, 0u) + __dp4a(w.w, 0x01010101u, 0u);
        }
    } else {
        for (uint32_t i = i0; i < i1; i++) sum += flags[i];
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) {
        const unsigned long long v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0ull;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned long long total = s_part[1023], before = s_part[threadIdx.x] - sum;
    if (threadIdx.x == 0) { bounds[0] = 0; bounds[1] = n; }   // shard 0 starts at 0, the last one ends at n; also the answer when nothing is flagged
    __syncthreads();
    // boundary b (b = shard, shard + 1) = the cell in which the running weight passes total * b / count; whoever holds it writes it
    for (uint32_t b = 0; b < 2u; b++) {
        const uint32_t which = shard + b;
        if (which == 0u || which == count || total == 0ull) continue;
        const unsigned long long target = frac ? min(total - 1ull, (unsigned long long) ((double) total * (double) fminf(fmaxf(frac[which], 0.0f), 1.0f)))
                                               : total * which / count;
        if (target >= before && target < before + sum) {
            unsigned long long seen = before;
            for (uint32_t i = i0; i < i1; i++) {
                seen += flags[i];
                if (seen > target) { bounds[b] = i; break; }
            }
        }
    }
}
__global__ void __launch_bounds__(256) k_take_shard(const float* __restrict__ in_vox, float* __restrict__ out_vox, DevState* st, int level,
                                                    uint32_t shard, uint32_t count, uint32_t* __restrict__ range_out,
                                                    const uint32_t* __restrict__ vp_in, uint32_t* __restrict__ vp_out,
                                                    const uint32_t* __restrict__ bounds /* precomputed [lo, hi), or null: equal parts */) {
    const uint32_t n = st->level_count[level];
    const uint32_t lo = bounds ? min(bounds[0], n) : (uint32_t) ((uint64_t) n * shard / count);
    const uint32_t hi = bounds ? min(max(bounds[1], lo), n) : (uint32_t) ((uint64_t) n * (shard + 1) / count);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 3u * (hi - lo); i += gridDim.x * blockDim.x) out_vox[i] = in_vox[3 * (size_t) lo + i];
    if (vp_in)   // list record (= parent) indices travel with the voxels
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < hi - lo; i += gridDim.x * blockDim.x) vp_out[i] = vp_in[lo + i];
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { range_out[0] = lo; range_out[1] = hi; range_out[2] = n; }
}
// all blocks must have read level_count before it is overwritten: done by a second, tiny launch
__global__ void k_set_level_count(DevState* st, int level, const uint32_t* __restrict__ range) { st->level_count[level] = range[1] - range[0]; }

// Sender side of the gather: vertex ids become global (offset of this shard's vertices in the merged table) and a
// triangle dropped by the finite filter is marked with id 0xFFFFFFFF in its first slot.
__global__ void __launch_bounds__(256) k_shard_prepare_send(DevState* st, uint32_t* __restrict__ tri_uid, const uint32_t* __restrict__ tri_valid_bits,
                                                            uint32_t vertex_offset) {
    const uint32_t T = st->n_tris_raw;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
        const bool valid = (tri_valid_bits[t >> 5] >> (t & 31u)) & 1u;
        tri_uid[3 * (size_t) t] = valid ? tri_uid[3 * (size_t) t] + vertex_offset : 0xFFFFFFFFu;
        tri_uid[3 * (size_t) t + 1] += vertex_offset;
        tri_uid[3 * (size_t) t + 2] += vertex_offset;
    }
}
// Root side: first-occurrence slots and validity bits over the merged triangle list (what k_orient does for one shard).
__global__ void __launch_bounds__(256) k_first_slot_merged(DevState* st, uint32_t* __restrict__ tri_uid, uint32_t* __restrict__ first_slot,
                                                           uint32_t* __restrict__ tri_valid_bits, uint32_t own_tris) {
    const uint32_t T = st->n_tris_raw;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t t0 = warp_id << 5; t0 < T; t0 += warps_total << 5) {
        const uint32_t t = t0 + lane;
        bool valid = false;
        if (t < T) {
            const uint32_t u0 = tri_uid[3 * (size_t) t];
            // the root's own triangles (t < own_tris) still carry their local validity bits, not the marker
            valid = t < own_tris ? ((tri_valid_bits[t >> 5] >> (t & 31u)) & 1u) : (u0 != 0xFFFFFFFFu);
            if (valid) {
                atomicMin(first_slot + u0, 3u * t);
                atomicMin(first_slot + tri_uid[3 * (size_t) t + 1], 3u * t + 1u);
                atomicMin(first_slot + tri_uid[3 * (size_t) t + 2], 3u * t + 2u);
            }
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) tri_valid_bits[t0 >> 5] = bits;
    }
}

// ---- distributed weld: local weld per shard, boundary keys resolved on rank 0, concatenation -------
// Shards DO share vertices: every edge of the interface layer is meshed by both neighbours, and any two vertices with the
// same quantised key are one vertex in the reference's weld (src/cuda/mod.rs:270-277).  Global first-occurrence order is
// shard-major (the triangle lists are concatenated in shard order), so a key's owner is its copy in the LOWEST shard, at
// that shard's local position; every other copy is removed from its shard and its index re-mapped to the owner.  With the
// welded local lists concatenated (offset Voff[s] = vertices of the shards before s), a kept vertex's global id is its
// concatenated position minus the number of removed vertices before it - one prefix pop-count over a removal bitmap.
// scratch words: [0] min ordered x, [1] max ordered x, [2] non-finite vertices, [3] boundary candidates
__global__ void k_shard_scratch_init(uint32_t* scratch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { scratch[0] = 0x7fffffffu; scratch[1] = 0x80000000u; scratch[2] = 0; scratch[3] = 0; }
}
__global__ void __launch_bounds__(256) k_shard_xrange(DevState* st, const float* __restrict__ out_pos, uint32_t* scratch) {
    const uint32_t n = st->n_verts_out;
    int lo = 0x7fffffff, hi = (int) 0x80000000;
    uint32_t bad = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = out_pos[3 * (size_t) i], y = out_pos[3 * (size_t) i + 1], z = out_pos[3 * (size_t) i + 2];
        // the range is over the x every vertex's KEY stands for: x itself, 0 for NaN, the far ends for +-inf (src/cuda/mod.rs:270: NaN -> 0,
        // `as i64` saturates) - a vertex of another shard with the same key has its x in this range whatever this vertex's y and z are
        const float xr = (x == x) ? fminf(fmaxf(x, -FLT_MAX), FLT_MAX) : 0.0f;
        lo = min(lo, f2ord(xr)); hi = max(hi, f2ord(xr));
        if (!(fabsf(x) <= FLT_MAX && fabsf(y) <= FLT_MAX && fabsf(z) <= FLT_MAX)) bad++;
    }
    lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi); bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31u) == 0) {
        atomicMin(reinterpret_cast<int*>(scratch), lo); atomicMax(reinterpret_cast<int*>(scratch) + 1, hi);
        if (bad) atomicAdd(scratch + 2, bad);
    }
}
// A vertex of this shard can share its key with a vertex of another shard only if its x lies within that shard's x range
// (equal keys: |x1 - x2| < 1.1e-5; the intervals handed in are widened by 1e-4).  Rows: kx, ky, kz, shard << 24 | local index.
struct ShardIntervals { float lo[32], hi[32]; uint32_t count; };
__global__ void __launch_bounds__(256) k_shard_boundary_keys(DevState* st, const float* __restrict__ out_pos, ShardIntervals iv, uint32_t shard,
                                                             uint4* __restrict__ out_rows, uint32_t cap_rows, uint32_t* scratch) {
    const uint32_t n = st->n_verts_out;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = out_pos[3 * (size_t) i];
        bool cand = false;
        for (uint32_t q = 0; q < iv.count; q++) cand = cand || (x >= iv.lo[q] && x <= iv.hi[q]);
        if (!cand) continue;
        const uint32_t slot = atomicAdd(scratch + 3, 1u);
        if (slot < cap_rows)
            out_rows[slot] = make_uint4(weld_key_component(x), weld_key_component(out_pos[3 * (size_t) i + 1]), weld_key_component(out_pos[3 * (size_t) i + 2]),
                                        (shard << 24) | i);
    }
}
struct ShardOffsets { uint32_t voff[33]; uint32_t poff[33]; uint32_t count; };   // vertices / duplicate pairs before each shard
// rank 0, pass 1: every row enters the key table; the entry keeps the smallest (shard, index) = the key's owner
__global__ void __launch_bounds__(256) k_res_insert(const uint4* __restrict__ rows, uint32_t total, uint4* table, uint32_t table_mask, uint32_t* __restrict__ rref,
                                                    uint32_t* err) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint4 r = rows[i];
        bool won;
        const uint32_t pos = hash_find_or_insert(table, table_mask, r.x, r.y, r.z, r.w, &won);
        rref[i] = pos;
        if (pos == 0xFFFFFFFFu) { atomicAdd(err, 1u); continue; }
        if (!won) atomicMin(reinterpret_cast<uint32_t*>(table + pos) + 3, r.w);
    }
}
// pass 2: a row that is not its key's owner is a duplicate: mark it in the removal bitmap (concatenated vertex space), count it
__global__ void __launch_bounds__(256) k_res_mark(const uint4* __restrict__ rows, uint32_t total, const uint4* __restrict__ table, const uint32_t* __restrict__ rref,
                                                  ShardOffsets so, uint32_t* __restrict__ bitmap, uint32_t* __restrict__ dups /* [shards] */) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t me = rows[i].w, pos = rref[i];
        if (pos == 0xFFFFFFFFu) continue;
        if (reinterpret_cast<const uint32_t*>(table + pos)[3] == me) continue;
        const uint32_t s = me >> 24, c = so.voff[s] + (me & 0xFFFFFFu);
        atomicOr(bitmap + (c >> 5), 1u << (c & 31u));
        atomicAdd(dups + s, 1u);
    }
}
// single-block exclusive prefix pop-count (the bitmaps here are ~1 MB)
__global__ void __launch_bounds__(1024) k_scan_bits_1block(const uint32_t* __restrict__ bits, uint32_t* __restrict__ word_prefix, uint32_t nwords) {
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (nwords + 1023u) / 1024u, w0 = threadIdx.x * per, w1 = min(w0 + per, nwords);
    uint32_t sum = 0;
    for (uint32_t w = w0; w < w1; w++) sum += __popc(bits[w]);
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) {
        const uint32_t v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;
    for (uint32_t w = w0; w < w1; w++) { word_prefix[w] = run; run += __popc(bits[w]); }
}
// pass 3: (local index, global id of the owner) for every duplicate, grouped by shard; global offset of every shard
__global__ void __launch_bounds__(256) k_res_pairs(const uint4* __restrict__ rows, uint32_t total, const uint4* __restrict__ table, const uint32_t* __restrict__ rref,
                                                   ShardOffsets so, const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                                                   uint32_t* __restrict__ cursor /* [shards], zeroed */, uint2* __restrict__ pairs, uint32_t* __restrict__ goff /* [shards] */) {
    if (blockIdx.x == 0 && threadIdx.x < so.count) goff[threadIdx.x] = so.voff[threadIdx.x] - bit_rank(bitmap, prefix, so.voff[threadIdx.x]);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t me = rows[i].w, pos = rref[i];
        if (pos == 0xFFFFFFFFu) continue;
        const uint32_t owner = reinterpret_cast<const uint32_t*>(table + pos)[3];
        if (owner == me) continue;
        const uint32_t s = me >> 24, oc = so.voff[owner >> 24] + (owner & 0xFFFFFFu);
        const uint32_t slot = so.poff[s] + atomicAdd(cursor + s, 1u);
        pairs[slot] = make_uint2(me & 0xFFFFFFu, oc - bit_rank(bitmap, prefix, oc));   // the owner itself is never removed
    }
}
// rank 0, after the welded shards (local indices, duplicates included) have arrived at their concatenated offsets:
// owner ids of the duplicates into a dense map over the concatenated vertex space ...
__global__ void __launch_bounds__(256) k_fix_scatter(const uint2* __restrict__ pairs, ShardOffsets so, uint32_t* __restrict__ remap) {
    const uint32_t total = so.poff[so.count];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t s = 0;
        while (s + 1 < so.count && i >= so.poff[s + 1]) s++;
        remap[so.voff[s] + pairs[i].x] = pairs[i].y;
    }
}
// ... kept vertices move up into the other output set (stable) ...
__global__ void __launch_bounds__(256) k_fix_vertices(uint32_t total_v, const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                                                      const float* __restrict__ in_pos, const float* __restrict__ in_nrm, float* __restrict__ out_pos,
                                                      float* __restrict__ out_nrm) {
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total_v; c += gridDim.x * blockDim.x) {
        if ((bitmap[c >> 5] >> (c & 31u)) & 1u) continue;
        const uint32_t o = c - bit_rank(bitmap, prefix, c);
#pragma unroll
        for (int k = 0; k < 3; k++) { out_pos[3 * (size_t) o + k] = in_pos[3 * (size_t) c + k]; out_nrm[3 * (size_t) o + k] = in_nrm[3 * (size_t) c + k]; }
    }
}
// ... and every index becomes global: shard-local index -> concatenated position -> owner id or own rank
struct ShardTriOffsets { uint32_t toff[33]; uint32_t count; };
__global__ void __launch_bounds__(256) k_fix_indices(ShardOffsets so, ShardTriOffsets to, const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                                                     const uint32_t* __restrict__ remap, uint32_t* __restrict__ idx) {
    const uint32_t n3 = 3u * to.toff[to.count];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n3; j += gridDim.x * blockDim.x) {
        const uint32_t t = j / 3u;
        uint32_t s = 0;
        while (s + 1 < to.count && t >= to.toff[s + 1]) s++;
        const uint32_t c = so.voff[s] + idx[j];
        idx[j] = ((bitmap[c >> 5] >> (c & 31u)) & 1u) ? remap[c] : c - bit_rank(bitmap, prefix, c);
    }
}

// ---- peer exchange: the distributed weld driven from the device (one process per GPU, peer-mapped memory on rank 0) ------------
// Rank 0 owns a control block, one slot of key rows per rank, the duplicate-pair lists and the output set the merged mesh is
// assembled in; every other rank maps them (CUDA IPC) and reads / writes them with plain loads and stores over NVLink.  A step
// never touches the host between its first and its last kernel:
//   P0  every rank : its shard, welded locally; header {V, T, x range, errors} -> rank 0's control block; flag A
//   P1  every rank : [all flags A] its welded vertices inside another shard's x range -> key rows in its slot on rank 0; flag B
//   P2  rank 0     : [all flags B] owner per key (lowest shard), removal bitmap + prefix over the concatenated vertex lists,
//                    per-shard global offsets and (local index, owner's global id) pairs; flag C
//   P3  every rank : [flag C] drops its own duplicates, makes its indices global and STORES its rows at their final offsets -
//                    straight into rank 0's output set over NVLink, or into its own second output set (host delivery: every
//                    rank then copies its part over its own PCIe link); flag D
//   P4  rank 0     : [all flags D] the merged mesh is complete (byte-identical to the single-GPU mesh).
// Flags carry the step's epoch (monotonic), the payload of step e lives in slot e & 1, so a fast rank may run one step ahead.
// Each flag is set by a one-thread kernel AFTER the kernels that wrote the payload (stream order) with a system-scope fence; the
// waiters are one-block kernels that spin on the flags (ld.acquire.sys).  With fewer GPUs than ranks (tests) the phases are
// issued one after the other with host synchronisation in between and no waiter is launched.
#define SDM_PEER_MAX 32
struct PeerHdr { uint32_t V, T, K, err; float min_x, max_x; uint32_t pad[2]; };
struct PeerCtl {
    uint32_t flagA[SDM_PEER_MAX], flagB[SDM_PEER_MAX], flagD[SDM_PEER_MAX];
    uint32_t flagC, pad0[31];
    PeerHdr hdr[2][SDM_PEER_MAX];
    uint32_t voff[2][SDM_PEER_MAX + 1], toff[2][SDM_PEER_MAX + 1], goff[2][SDM_PEER_MAX + 1], poff[2][SDM_PEER_MAX + 1];
    uint32_t dups[2][SDM_PEER_MAX], cursor[2][SDM_PEER_MAX];
    uint32_t status[2], total_rows[2];
};
struct PeerLocal { uint32_t goff, toff, kept, T, status, total_V, total_T, pad; };   // what a rank's host reads after a step

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__global__ void k_peer_scan_reset(DevState* st, int which) {
    if (threadIdx.x == 0 && blockIdx.x == 0) st->ticket[(which & 1) == 0 ? TK_SCAN_FIRST : TK_SCAN_TRI] = 0;
}
// one block; thread i waits for flags[i] to reach `epoch`.  A peer that never arrives (its process died) must not hang this GPU:
// after 20 s the waiter gives up and marks the step failed.
__global__ void k_peer_wait(const uint32_t* flags, uint32_t count, uint32_t epoch, DevState* st) {
    if (threadIdx.x < count) {
        unsigned long long t0 = 0;
        uint32_t spins = 0;
        while ((int32_t) (ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
            __nanosleep(200);
            if ((++spins & 0xFFFu) == 0u) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > 20000000000ull) { atomicOr(&st->error_flags, ERR_PEER_TIMEOUT); break; }
            }
        }
    }
}
__global__ void k_peer_set_flag(uint32_t* flag, uint32_t epoch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { __threadfence_system(); st_release_sys(flag, epoch); }
}
// P0: header of this rank's welded shard (scratch: k_shard_xrange's {min ordered x, max ordered x, non-finite count})
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_peer_mark_start(unsigned long long* t0) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *t0 = global_timer_ns();
}
// Load balancing from measurement.  The shards' compute times of the previous step (in their headers; stable and the same for every
// rank by the time a step starts) and the cumulative weight fractions that step used give a piecewise-constant cost density over the
// weight axis; the new boundaries cut the cumulative cost into equal parts, half-way damped.  Every rank runs this on the same
// inputs and gets the same fractions, so the shards still tile the list; any split gives the same mesh.
__global__ void k_peer_rebalance(const PeerCtl* ctl, uint32_t prev_parity, uint32_t world, const float* __restrict__ frac_prev, float* __restrict__ frac_new, int have_prev) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    bool ok = have_prev != 0 && ctl->status[prev_parity] == 0u;
    float t[SDM_PEER_MAX], total = 0.0f;
    for (uint32_t r = 0; r < world; r++) {
        t[r] = (float) ctl->hdr[prev_parity][r].pad[0];
        ok = ok && t[r] > 0.0f && frac_prev[r + 1] > frac_prev[r];
        total += t[r];
    }
    frac_new[0] = 0.0f; frac_new[world] = 1.0f;
    if (!ok) { for (uint32_t b = 1; b < world; b++) frac_new[b] = (float) b / (float) world; return; }
    float c0 = 0.0f;
    uint32_t r = 0;
    for (uint32_t b = 1; b < world; b++) {
        const float target = total * (float) b / (float) world;
        while (r + 1 < world && c0 + t[r] < target) { c0 += t[r]; r++; }
        const float w = frac_prev[r] + (target - c0) / t[r] * (frac_prev[r + 1] - frac_prev[r]);
        float f = 0.5f * frac_prev[b] + 0.5f * w;
        f = fmaxf(f, frac_new[b - 1] + 1e-4f);
        frac_new[b] = fminf(f, 1.0f - 1e-4f * (float) (world - b));
    }
}
__global__ void k_peer_publish_header(DevState* st, const uint32_t* __restrict__ scratch, PeerCtl* ctl, uint32_t rank, uint32_t parity, const unsigned long long* t0) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    PeerHdr hd;
    hd.V = st->n_verts_out; hd.T = st->n_tris_out; hd.K = 0;
    hd.err = st->error_flags | (st->n_verts_out >= (1u << 24) ? 0x200u : 0u);   // 24-bit local indices in the key rows
    hd.min_x = ord2f((int) scratch[0]); hd.max_x = ord2f((int) scratch[1]);
    if (scratch[0] == 0x7fffffffu) { hd.min_x = 1.0f; hd.max_x = 0.0f; }   // no finite vertex
    hd.pad[0] = (uint32_t) min((global_timer_ns() - *t0) / 1000ull, 0xFFFFFFFFull);   // this shard's compute time in microseconds (k_peer_rebalance)
    hd.pad[1] = 0;
    ctl->hdr[parity][rank] = hd;
}
// P1: key rows of the welded vertices that lie inside another shard's x range (equal keys are < 1.1e-5 apart; the ranges are
// widened by 1e-4, or by four units of the key grid's float spacing where coordinates are large) -> this rank's slot on rank 0
__global__ void __launch_bounds__(256) k_peer_rows(DevState* st, const float* __restrict__ out_pos, const PeerCtl* ctl, uint32_t rank, uint32_t world,
                                                   uint32_t parity, uint4* __restrict__ slot_rows, uint32_t cap_rows, uint32_t* counter) {
    __shared__ float s_lo[SDM_PEER_MAX], s_hi[SDM_PEER_MAX];
    if (threadIdx.x < world) {
        const PeerHdr hd = ctl->hdr[parity][threadIdx.x];
        const float m = fmaxf(fabsf(hd.min_x), fabsf(hd.max_x));
        const float wd = fmaxf(1e-4f, m * 4.8e-7f);
        const bool use = threadIdx.x != rank && hd.min_x <= hd.max_x;
        s_lo[threadIdx.x] = use ? hd.min_x - wd : 1.0f;
        s_hi[threadIdx.x] = use ? hd.max_x + wd : 0.0f;
    }
    __syncthreads();
    const uint32_t n = st->n_verts_out;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = out_pos[3 * (size_t) i], y = out_pos[3 * (size_t) i + 1], z = out_pos[3 * (size_t) i + 2];
        // a vertex with a non-finite coordinate is in no x range but still has a key (NaN -> 0, src/cuda/mod.rs:270): always a candidate
        bool cand = !(fabsf(x) <= FLT_MAX && fabsf(y) <= FLT_MAX && fabsf(z) <= FLT_MAX);
        for (uint32_t q = 0; q < world; q++) cand = cand || (x >= s_lo[q] && x <= s_hi[q]);
        if (!cand) continue;
        const uint32_t slot = atomicAdd(counter, 1u);
        if (slot < cap_rows) slot_rows[slot] = make_uint4(weld_key_component(x), weld_key_component(y), weld_key_component(z), (rank << 24) | i);
    }
}
__global__ void k_peer_publish_rows(const uint32_t* counter, PeerCtl* ctl, uint32_t rank, uint32_t parity, uint32_t cap_rows) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t k = *counter;
    ctl->hdr[parity][rank].K = min(k, cap_rows);
    if (k > cap_rows) ctl->hdr[parity][rank].err |= 0x400u;   // more boundary candidates than the slot holds
}
// P2 (rank 0), step 1: offsets of the concatenated vertex / triangle lists; any rank's error fails the step for everybody
__global__ void k_peer_root_offsets(PeerCtl* ctl, uint32_t world, uint32_t parity, uint32_t cap_vertices, uint32_t cap_triangles) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t v = 0, t = 0, k = 0, err = 0;
    for (uint32_t r = 0; r < world; r++) {
        const PeerHdr hd = ctl->hdr[parity][r];
        ctl->voff[parity][r] = v; ctl->toff[parity][r] = t;
        if ((uint64_t) v + hd.V > 0xFFFFFFFFull || (uint64_t) t + hd.T > 0xFFFFFFFFull) err |= 0x800u;
        v += hd.V; t += hd.T; k += hd.K; err |= hd.err;
        ctl->dups[parity][r] = 0; ctl->cursor[parity][r] = 0;
    }
    ctl->voff[parity][world] = v; ctl->toff[parity][world] = t;
    if (v > cap_vertices || t > cap_triangles) err |= 0x800u;   // the merged mesh does not fit in rank 0's output set
    ctl->status[parity] = err;
    ctl->total_rows[parity] = k;
}
// rank 0's key table for one step: sized on the device from the rows that actually arrived (load factor <= 1/2)
__device__ __forceinline__ uint32_t peer_table_size(uint32_t total_rows, uint32_t max_entries) {
    uint32_t s = 1024;
    while ((uint64_t) s < (uint64_t) total_rows * 2u && s < max_entries) s <<= 1;
    return s;
}
__global__ void __launch_bounds__(256) k_peer_clear_table(const PeerCtl* ctl, uint32_t parity, uint4* __restrict__ table, uint32_t max_entries) {
    const uint32_t n = peer_table_size(ctl->total_rows[parity], max_entries);
    const uint4 empty = make_uint4(SDM_HASH_EMPTY, SDM_HASH_EMPTY, SDM_HASH_EMPTY, SDM_HASH_EMPTY);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) table[i] = empty;
}
// rows live in per-rank slots of cap_rows entries: row g of the flattened space is rows[g] if (g % cap_rows) < K of rank g / cap_rows
__device__ __forceinline__ bool peer_row(const PeerCtl* ctl, uint32_t parity, uint32_t cap_rows, uint32_t g) { return (g % cap_rows) < ctl->hdr[parity][g / cap_rows].K; }
__global__ void __launch_bounds__(256) k_peer_res_insert(const PeerCtl* ctl, uint32_t world, uint32_t parity, uint32_t cap_rows, const uint4* __restrict__ rows, uint4* table,
                                                         uint32_t max_entries, uint32_t* __restrict__ rref, uint32_t* status) {
    if (*status) return;
    const uint32_t table_mask = peer_table_size(ctl->total_rows[parity], max_entries) - 1u;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < world * cap_rows; g += gridDim.x * blockDim.x) {
        if (!peer_row(ctl, parity, cap_rows, g)) continue;
        const uint4 r = rows[g];
        bool won;
        const uint32_t pos = hash_find_or_insert(table, table_mask, r.x, r.y, r.z, r.w, &won);
        rref[g] = pos;
        if (pos == 0xFFFFFFFFu) { atomicOr(status, 0x1000u); continue; }
        if (!won) atomicMin(reinterpret_cast<uint32_t*>(table + pos) + 3, r.w);
    }
}
// a row that is not its key's owner is a duplicate: bit in the removal bitmap (concatenated vertex space), count per shard
__global__ void __launch_bounds__(256) k_peer_res_mark(PeerCtl* ctl, uint32_t world, uint32_t parity, uint32_t cap_rows, const uint4* __restrict__ rows, const uint4* __restrict__ table,
                                                       const uint32_t* __restrict__ rref, uint32_t* __restrict__ bitmap) {
    if (ctl->status[parity]) return;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < world * cap_rows; g += gridDim.x * blockDim.x) {
        if (!peer_row(ctl, parity, cap_rows, g)) continue;
        const uint32_t me = rows[g].w, pos = rref[g];
        if (pos == 0xFFFFFFFFu || reinterpret_cast<const uint32_t*>(table + pos)[3] == me) continue;
        const uint32_t s = me >> 24, c = ctl->voff[parity][s] + (me & 0xFFFFFFu);
        atomicOr(bitmap + (c >> 5), 1u << (c & 31u));
        atomicAdd(&ctl->dups[parity][s], 1u);
    }
}
__global__ void k_peer_root_goff(PeerCtl* ctl, uint32_t world, uint32_t parity) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t g = 0, p = 0;
    for (uint32_t r = 0; r < world; r++) {
        ctl->goff[parity][r] = g; ctl->poff[parity][r] = p;
        g += ctl->hdr[parity][r].V - ctl->dups[parity][r]; p += ctl->dups[parity][r];
    }
    ctl->goff[parity][world] = g; ctl->poff[parity][world] = p;
}
// (local index, global id of the owner) for every duplicate, grouped by shard
__global__ void __launch_bounds__(256) k_peer_res_pairs(PeerCtl* ctl, uint32_t world, uint32_t parity, uint32_t cap_rows, const uint4* __restrict__ rows, const uint4* __restrict__ table,
                                                        const uint32_t* __restrict__ rref, const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                                                        uint2* __restrict__ pairs) {
    if (ctl->status[parity]) return;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < world * cap_rows; g += gridDim.x * blockDim.x) {
        if (!peer_row(ctl, parity, cap_rows, g)) continue;
        const uint32_t me = rows[g].w, pos = rref[g];
        if (pos == 0xFFFFFFFFu) continue;
        const uint32_t owner = reinterpret_cast<const uint32_t*>(table + pos)[3];
        if (owner == me) continue;
        const uint32_t s = me >> 24, oc = ctl->voff[parity][owner >> 24] + (owner & 0xFFFFFFu);
        const uint32_t slot = ctl->poff[parity][s] + atomicAdd(&ctl->cursor[parity][s], 1u);
        pairs[slot] = make_uint2(me & 0xFFFFFFu, oc - bit_rank(bitmap, prefix, oc));   // the owner itself is never removed
    }
}
// P3: this rank's duplicates into its own removal bitmap + owner map (local vertex space)
__global__ void __launch_bounds__(256) k_peer_apply_mark(const PeerCtl* ctl, uint32_t rank, uint32_t parity, const uint2* __restrict__ pairs, uint32_t* __restrict__ bitmap,
                                                         uint32_t* __restrict__ remap, PeerLocal* local) {
    const uint32_t p0 = ctl->poff[parity][rank], np = ctl->poff[parity][rank + 1] - p0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        local->goff = ctl->goff[parity][rank]; local->toff = ctl->toff[parity][rank]; local->kept = ctl->hdr[parity][rank].V - np;
        local->T = ctl->hdr[parity][rank].T; local->status = ctl->status[parity];
        local->total_V = 0; local->total_T = 0;
    }
    if (ctl->status[parity]) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const uint2 pr = pairs[p0 + i];
        atomicOr(bitmap + (pr.x >> 5), 1u << (pr.x & 31u));
        remap[pr.x] = pr.y;
    }
}
// kept vertices to their global rows (dst: rank 0's output set over NVLink, or this rank's own second set)
__global__ void __launch_bounds__(256) k_peer_apply_vertices(const PeerCtl* ctl, uint32_t rank, uint32_t parity, const uint32_t* __restrict__ bitmap,
                                                             const uint32_t* __restrict__ prefix, const float* __restrict__ in_pos, const float* __restrict__ in_nrm,
                                                             float* __restrict__ dst_pos, float* __restrict__ dst_nrm, uint32_t dst_is_global) {
    if (ctl->status[parity]) return;
    const uint32_t V = ctl->hdr[parity][rank].V;
    const uint32_t base = dst_is_global ? ctl->goff[parity][rank] : 0u;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < V; c += gridDim.x * blockDim.x) {
        if ((bitmap[c >> 5] >> (c & 31u)) & 1u) continue;
        const size_t o = (size_t) base + c - bit_rank(bitmap, prefix, c);
#pragma unroll
        for (int k = 0; k < 3; k++) { dst_pos[3 * o + k] = in_pos[3 * (size_t) c + k]; dst_nrm[3 * o + k] = in_nrm[3 * (size_t) c + k]; }
    }
}
// every index becomes global: own kept vertex -> goff + its rank among the kept ones, duplicate -> its owner's global id
__global__ void __launch_bounds__(256) k_peer_apply_indices(const PeerCtl* ctl, uint32_t rank, uint32_t parity, const uint32_t* __restrict__ bitmap,
                                                            const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ remap, const uint32_t* __restrict__ in_idx,
                                                            uint32_t* __restrict__ dst_idx, uint32_t dst_is_global) {
    if (ctl->status[parity]) return;
    const uint32_t n3 = 3u * ctl->hdr[parity][rank].T, goff = ctl->goff[parity][rank];
    const size_t base = dst_is_global ? 3 * (size_t) ctl->toff[parity][rank] : 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n3; j += gridDim.x * blockDim.x) {
        const uint32_t c = in_idx[j];
        dst_idx[base + j] = ((bitmap[c >> 5] >> (c & 31u)) & 1u) ? remap[c] : goff + c - bit_rank(bitmap, prefix, c);
    }
}
// deliver = 0: this rank's final rows (compacted in its own second output set) go to rank 0's output set at their global offsets as
// plain coalesced word copies - 128 contiguous bytes per warp store, the packet size NVLink likes (12-byte stores scattered by
// k_peer_apply_* straight into the peer took three times as long).  which: 0 positions, 1 normals, 2 indices.
__global__ void __launch_bounds__(256) k_peer_push(const PeerCtl* ctl, uint32_t rank, uint32_t parity, const uint32_t* __restrict__ src_pos, const uint32_t* __restrict__ src_nrm,
                                                   const uint32_t* __restrict__ src_idx, uint32_t* __restrict__ dst_pos, uint32_t* __restrict__ dst_nrm, uint32_t* __restrict__ dst_idx) {
    if (ctl->status[parity]) return;
    const size_t nv = 3 * (size_t) (ctl->goff[parity][rank + 1] - ctl->goff[parity][rank]), nt = 3 * (size_t) ctl->hdr[parity][rank].T;
    const size_t ov = 3 * (size_t) ctl->goff[parity][rank], ot = 3 * (size_t) ctl->toff[parity][rank];
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t) gridDim.x * blockDim.x;
    // 16-byte STORES (what crosses NVLink): a few head words bring the destination to a 16-byte boundary (row offsets are multiples
    // of 12 bytes), the body gathers four source words - whatever their alignment - per store
    auto copy = [&](const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n) {
        const size_t head = min(n, (size_t) ((16u - (uint32_t) (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) >> 2);
        for (size_t i = tid; i < head; i += stride) dst[i] = src[i];
        const size_t n4 = (n - head) >> 2;
        uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst + head);
        const uint32_t* __restrict__ sb = src + head;
        for (size_t i = tid; i < n4; i += stride) d4[i] = make_uint4(sb[4 * i], sb[4 * i + 1], sb[4 * i + 2], sb[4 * i + 3]);
        for (size_t i = head + (n4 << 2) + tid; i < n; i += stride) dst[i] = src[i];
    };
    copy(src_pos, dst_pos + ov, nv);
    copy(src_nrm, dst_nrm + ov, nv);
    copy(src_idx, dst_idx + ot, nt);
}
// totals for the host of rank 0 (and of every rank, for bookkeeping)
__global__ void k_peer_totals(const PeerCtl* ctl, uint32_t world, uint32_t parity, PeerLocal* local) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    local->total_V = ctl->goff[parity][world]; local->total_T = ctl->toff[parity][world]; local->status = ctl->status[parity];
}

// ---- test / probe kernels -------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_eval_sdf(const uint4* __restrict__ scene, const float* __restrict__ pts, uint32_t n, float* __restrict__ out,
                                                  MaskGrid grid) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5, warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t i0 = warp_id << 5; i0 < n; i0 += warps_total << 5) {
        const uint32_t i = i0 + lane;
        const bool active = i < n;
        float x = 0.f, y = 0.f, z = 0.f;
        if (active) { x = pts[3 * (size_t) i]; y = pts[3 * (size_t) i + 1]; z = pts[3 * (size_t) i + 2]; }
        tile_mask_from_point(grid, sc, active, x, y, z);
        if (active) out[i] = eval_scene1(sc, x, y, z);
    }
}
__global__ void __launch_bounds__(128) k_eval_normal(const uint4* __restrict__ scene, const float* __restrict__ pts, uint32_t n, float* __restrict__ out,
                                                     MaskGrid grid) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5, warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t i0 = warp_id << 5; i0 < n; i0 += warps_total << 5) {
        const uint32_t i = i0 + lane;
        const bool active = i < n;
        float x = 0.f, y = 0.f, z = 0.f;
        if (active) { x = pts[3 * (size_t) i]; y = pts[3 * (size_t) i + 1]; z = pts[3 * (size_t) i + 2]; }
        tile_mask_from_point(grid, sc, active, x, y, z);
        if (active) {
            float nx, ny, nz;
            empirical_normal(sc, x, y, z, nx, ny, nz);
            out[3 * (size_t) i] = nx; out[3 * (size_t) i + 1] = ny; out[3 * (size_t) i + 2] = nz;
        }
    }
}
__global__ void __launch_bounds__(128) k_eval_project(const uint4* __restrict__ scene, const float* __restrict__ pts, uint32_t n,
                                                      float* __restrict__ out, uint32_t* __restrict__ iters, MaskGrid grid) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5, warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t i0 = warp_id << 5; i0 < n; i0 += warps_total << 5) {
        const uint32_t i = i0 + lane;
        bool running = i < n;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (running) { gx = pts[3 * (size_t) i]; gy = pts[3 * (size_t) i + 1]; gz = pts[3 * (size_t) i + 2]; }
        uint32_t it = 0;
        bool collision = false;
        NewtonCycle cyc;
        cyc.start(gx, gy, gz);
        while (__any_sync(0xffffffffu, running)) {
            tile_mask_from_point(grid, sc, running, gx, gy, gz);
            if (running) {
                collision = newton_step(sc, gx, gy, gz);
                it++;
                if (!collision) cyc.observe(gx, gy, gz, it);
                running = !collision && it < cyc.stop_at;
            }
        }
        if (i < n) {
            out[3 * (size_t) i] = gx; out[3 * (size_t) i + 1] = gy; out[3 * (size_t) i + 2] = gz;
            if (iters) iters[i] = collision ? it : 10000u;   // the reference's iteration count
        }
    }
}

// Exhaustive / randomised comparison of the branch-free sqrt and division with sqrtf and `/` (sdm_selftest_math).
__global__ void __launch_bounds__(256) k_selftest_math(unsigned long long* __restrict__ out /* {sqrt mismatches, sqrt fallbacks, div mismatches, div fallbacks} */,
                                                       unsigned long long div_samples) {
    unsigned long long sm = 0, sf = 0, dm = 0, df = 0;
    const unsigned long long tid = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long i = tid; i < (1ull << 32); i += stride) {
        const float x = __uint_as_float((uint32_t) i);
        bool bad = false;
        const float a = sqrt_nobranch(x, bad);
        if (bad) { sf++; continue; }
        const float b = sqrtf(x);
        if (__float_as_uint(a) != __float_as_uint(b) && !(x == 0.0f)) sm++;
    }
    for (unsigned long long i = tid; i < div_samples; i += stride) {
        // k log-uniform in [1e-6, 1e6] plus the scenes' own 0.1 / 0.5; t = k * u, u in (0, 1]
        uint32_t r0 = (uint32_t) (i * 0x9E3779B97F4A7C15ull >> 32), r1 = (uint32_t) ((i ^ 0xD1B54A32D192ED03ull) * 0xBF58476D1CE4E5B9ull >> 32);
        r0 ^= r0 >> 15; r0 *= 0x2C1B3C6Du; r0 ^= r0 >> 12; r1 ^= r1 >> 13; r1 *= 0x297A2D39u; r1 ^= r1 >> 15;
        float k = (i & 3ull) == 0 ? 0.1f : ((i & 3ull) == 1 ? 0.5f : __uint_as_float(0x358637BDu + r0 % (0x49742400u - 0x358637BDu)));
        const float u = (float) (r1 >> 8) * (1.0f / 16777216.0f);
        const float t = (i & 4ull) ? k * u : __uint_as_float(r1 & 0x7FFFFFFFu);   // also arbitrary positive numerators
        if (!(t > 0.0f) || !(t <= 1e30f)) continue;
        if (t < SDM_DIV_GUARD_LO) { df++; continue; }
        const float y = div_prepare(k);
        const float a = div_nobranch(t, k, y);
        const float b = t / k;
        if (__float_as_uint(a) != __float_as_uint(b)) dm++;
    }
    atomicAdd(out + 0, sm); atomicAdd(out + 1, sf); atomicAdd(out + 2, dm); atomicAdd(out + 3, df);
}

// ---- primitive masks --------------------------------------------------------------------------------------
// For a cell with centre c and radius rho (circumsphere of the cell cube + slop, see host code), primitive i is dropped
// iff   d_i(c) - rho  >=  U_i + k_i + margin,   U_i = min over earlier primitives j < i (of the parent cell's mask, or
// all of them) of d_j(c) + rho.
// Why that is exact.  All primitive kinds admitted here (sphere, capsule, box) are 1-Lipschitz distance functions, so for
// every p within rho of c:  d_i(p) >= d_i(c) - rho  and  d_j(p) <= d_j(c) + rho.  The fold accumulator never exceeds the
// minimum of the distances folded so far (fminf(a,b) - h*h*h*k/6 <= fminf(a,b); signed_distance.cu:21-22), and a
// skipped primitive leaves it unchanged, so acc_i(p) <= U_i.  Hence d_i(p) - acc_i(p) >= k_i + margin: for smooth_min,
// k - |acc - d| <= 0 gives h = 0 and the result is fminf(acc, d) - 0 = acc bit for bit; for min, fminf(acc, d) = acc.
// The margin (1e-4) is orders of magnitude above the rounding error of the distances involved (|d| < ~10).
// With parent masks (coarse grid, 4x4x4 fine cells per coarse cell) only primitives of the parent's mask are tested:
// a primitive outside the parent's mask is already proven droppable on the parent's sphere, which contains the child's.
__global__ void __launch_bounds__(256) k_build_masks(const uint4* __restrict__ scene, uint32_t* __restrict__ out_masks, MaskGrid g,
                                                     const uint32_t* __restrict__ parent_masks, uint32_t parent_G, float rho,
                                                     uint8_t* __restrict__ out_maybe) {
    // the table is read through L1 (each lane reads a different record: from shared memory that would be a 16-way bank
    // conflict, and staging 64 KB per block would cap occupancy)
    const SceneHeader hdr = *reinterpret_cast<const SceneHeader*>(scene);
    SceneView sc;
    sc.runs = nullptr; sc.nruns = 0; sc.nprims = hdr.nprims; sc.wmask = nullptr; sc.W = 0; sc.tlist = nullptr; sc.tcount = nullptr;
    sc.prims = reinterpret_cast<const DevPrim*>(scene + 1 + hdr.nruns);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5, warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t ncells = g.G * g.G * g.G;
    const float inf = __int_as_float(0x7f800000);
    for (uint32_t cell = warp_id; cell < ncells; cell += warps_total) {
        const uint32_t iz = cell % g.G, iy = (cell / g.G) % g.G, ix = cell / (g.G * g.G);
        const float cx = g.ox + ((float) ix + 0.5f) * g.cell, cy = g.oy + ((float) iy + 0.5f) * g.cell, cz = g.oz + ((float) iz + 0.5f) * g.cell;
        const uint32_t* prow = nullptr;
        if (parent_masks) {
            const uint32_t f = g.G / parent_G;
            prow = parent_masks + ((size_t) (((ix / f) * parent_G + iy / f) * parent_G + iz / f)) * g.W;
        }
        float carry = inf;
        // lane w keeps parent word w and result word w: one coalesced load and one coalesced store per cell (32 words per
        // pass), and only the non-empty parent words are visited
        for (uint32_t w0 = 0; w0 < g.W; w0 += 32u) {
            const uint32_t wl = w0 + lane;
            const uint32_t pmw = wl < g.W ? (prow ? prow[wl] : 0xFFFFFFFFu) : 0u;
            uint32_t myword = 0;
            uint32_t nz = __ballot_sync(0xffffffffu, pmw != 0u);
            while (nz) {
                const uint32_t wi = (uint32_t) __ffs((int) nz) - 1u;
                nz &= nz - 1u;
                const uint32_t pm = __shfl_sync(0xffffffffu, pmw, wi);
                const uint32_t j = ((w0 + wi) << 5) + lane;
                const bool active = ((pm >> lane) & 1u) && j < sc.nprims;
                float d = inf, kk = 0.0f;
                if (active) {
                    const DevPrim c = sc.prims[j];
                    d = prim_distance_cull(c, cx, cy, cz);
                    kk = c.fold == SDM_FOLD_SMOOTH_MIN ? c.k : 0.0f;
                }
                const float ub = d + rho;
                float e = ub;   // inclusive prefix-min over lanes, then shifted to exclusive
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v = __shfl_up_sync(0xffffffffu, e, o);
                    if (lane >= (uint32_t) o) e = fminf(e, v);
                }
                const float total = __shfl_sync(0xffffffffu, e, 31);
                float excl = __shfl_up_sync(0xffffffffu, e, 1);
                if (lane == 0) excl = inf;
                const float U = fminf(carry, excl);
                const bool keep = active && !(d - rho >= U + kk + 1e-4f);   // NaN distance: keep
                const uint32_t word = __ballot_sync(0xffffffffu, keep);
                if (lane == wi) myword = word;
                carry = fminf(carry, total);
            }
            if (wl < g.W) out_masks[(size_t) cell * g.W + wl] = myword;
        }
        // Zero-crossing flag.  carry - rho = min_i d_i(c) (a primitive outside the parent's mask is never the nearest one).  For
        // every p within rho of c:  min_i d_i(p) - kmax <= sd(p) <= min_i d_i(p)  (smooth_min(a,b) <= min(a,b); the fold never
        // drops more than kmax below the minimum of the distances folded so far, see tile_refine) and |d_i(p) - d_i(c)| <= rho,
        // so  min_d - rho - kmax > 0  means sd > 0 on the whole cell and  min_d + rho < 0  means sd < 0 on the whole cell.
        if (out_maybe && lane == 0) {
            const float min_d = carry - rho;
            const bool empty = (min_d - rho - hdr.kmax > 1e-4f) || (min_d + rho < -1e-4f);
            out_maybe[cell] = empty ? 0 : 1;   // NaN: not empty
        }
    }
}

// Fine level of the mask build for tables of at most 1024 primitives (W <= 32): one LANE per cell.  A warp takes 32 of the 64
// fine cells of one coarse cell and walks that parent's candidates in index order; every record is a warp-uniform
// (broadcast) load and each lane runs the drop test above at its own cell centre with its own running U.  (With one warp per
// cell and one candidate per lane, every round gathered 32 different 64-byte records: the kernel spent its time in the
// load/store unit - 1.3 ms for 64^3 cells against ~0.1 ms this way.)  Kept bits go to a per-warp tile of rows in shared
// memory (row stride 33 words: lane-private rows, conflict-free) and are written out as whole rows.
__global__ void __launch_bounds__(256) k_build_masks_fine(const uint4* __restrict__ scene, uint32_t* __restrict__ out_masks, MaskGrid g,
                                                          const uint32_t* __restrict__ parent_masks, uint32_t parent_G, float rho,
                                                          uint8_t* __restrict__ out_maybe) {
    __shared__ uint32_t s_rows[8][32 * 33];
    const SceneHeader hdr = *reinterpret_cast<const SceneHeader*>(scene);
    const DevPrim* __restrict__ prims = reinterpret_cast<const DevPrim*>(scene + 1 + hdr.nruns);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5, warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t* rows = s_rows[warp];
    const uint32_t f = g.G / parent_G;                 // 4
    const uint32_t ntasks = parent_G * parent_G * parent_G * 2u;   // two half coarse cells (2 x 4 x 4 fine cells) each
    const float inf = __int_as_float(0x7f800000);
    for (uint32_t task = warp_id; task < ntasks; task += warps_total) {
        const uint32_t pc = task >> 1, half = task & 1u;
        const uint32_t pz = pc % parent_G, py = (pc / parent_G) % parent_G, px = pc / (parent_G * parent_G);
        const uint32_t ix = px * f + half * 2u + (lane >> 4), iy = py * f + ((lane >> 2) & 3u), iz = pz * f + (lane & 3u);
        const float cx = g.ox + ((float) ix + 0.5f) * g.cell, cy = g.oy + ((float) iy + 0.5f) * g.cell, cz = g.oz + ((float) iz + 0.5f) * g.cell;
        const uint32_t* __restrict__ prow = parent_masks + (size_t) pc * g.W;
        for (uint32_t w = 0; w < g.W; w++) rows[lane * 33u + w] = 0u;
        // acc = the scene fold at the cell centre over the primitives kept so far (a dropped primitive would not have changed it),
        // with the approximate distances: within ~1e-5 of the exact value.  The fold of 1-Lipschitz distances is 1-Lipschitz,
        // so acc + rho bounds the accumulator on the whole cell from above - a tighter U than min_j d_j(c) + rho wherever
        // primitives blend - and |acc_final| > rho means the SDF keeps its sign on the cell (zero-crossing flag below).
        float acc = SDM_MAX_POSITIVE_F32;
        for (uint32_t w = 0; w < g.W; w++) {
            uint32_t pm = __ldg(prow + w);   // warp-uniform
            uint32_t mine = 0;
            while (pm) {
                const uint32_t b = (uint32_t) __ffs((int) pm) - 1u;
                pm &= pm - 1u;
                const uint32_t j = (w << 5) + b;
                if (j >= hdr.nprims) break;
                const DevPrim c = prims[j];
                const float d = prim_distance_cull(c, cx, cy, cz);
                const float kk = c.fold == SDM_FOLD_SMOOTH_MIN ? c.k : 0.0f;
                if (!(d - rho >= (acc + rho) + kk + 1e-4f)) {   // NaN distance: keep
                    mine |= 1u << b;
                    acc = c.fold == SDM_FOLD_SMOOTH_MIN ? smooth_min_skip(acc, d, c.k) : fminf(acc, d);
                }
            }
            rows[lane * 33u + w] = mine;
        }
        __syncwarp();
        // whole rows out: row r belongs to the cell of lane r
        for (uint32_t r = 0; r < 32u; r++) {
            const uint32_t rx = px * f + half * 2u + (r >> 4), ry = py * f + ((r >> 2) & 3u), rz = pz * f + (r & 3u);
            const size_t cell = ((size_t) rx * g.G + ry) * g.G + rz;
            if (lane < g.W) out_masks[cell * g.W + lane] = rows[r * 33u + lane];
        }
        if (out_maybe) {   // zero-crossing flag: |sd(p) - sd(c)| <= rho on the cell, acc is sd(c) to ~1e-5
            const bool empty = fabsf(acc) > rho + 1e-4f;
            // non-zero = may contain the surface; the value is the number of primitives the cell keeps (1..255), which the multi-GPU
            // split uses as the cell's weight: the cost of a surface voxel grows with the length of its primitive list
            uint32_t kept = 0;
            for (uint32_t w = 0; w < g.W; w++) kept += __popc(rows[lane * 33u + w]);
            out_maybe[((size_t) ix * g.G + iy) * g.G + iz] = empty ? 0 : (uint8_t) max(1u, min(255u, kept));   // NaN: not empty
        }
        __syncwarp();
    }
}

}  // namespace sdm
