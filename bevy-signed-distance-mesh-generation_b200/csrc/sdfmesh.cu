// sdfmesh.cu - host side of libsdfmesh.so: the C ABI of include/sdfmesh.h over the kernels in sdm_kernels.cuh.
//
// Mirrors the reference's CudaHandler (src/cuda/mod.rs): grow-only device buffers owned by the handle
// (DynamicCudaSlice, :10-23), one device, one stream.  Unlike the reference nothing is compacted or welded on
// the host: the voxel list, the case indices, the vertex table and the index buffer all stay in HBM, and one
// remesh is enqueued as a fixed sequence of persistent kernels that read their sizes from device memory; the
// host synchronises once, at the end, to learn the counts.
//
// There is no CPU fallback: every compute entry point needs a CUDA device and fails with SDM_ERR_NO_DEVICE /
// SDM_ERR_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "sdm_kernels.cuh"
#include "sdm_render.cuh"

// Device-side clears go through a kernel (one engine for everything on the compute stream; the copy engines are left to the
// asynchronous mesh download).  `bytes` and `p` must be multiples of 4.
__global__ void __launch_bounds__(256) k_fill32(uint32_t* __restrict__ p, uint32_t value, size_t nwords) {
    const size_t n4 = nwords >> 2;
    const uint4 v4 = make_uint4(value, value, value, value);
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t) gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        for (size_t i = tid; i < n4; i += stride) reinterpret_cast<uint4*>(p)[i] = v4;
        for (size_t i = (n4 << 2) + tid; i < nwords; i += stride) p[i] = value;
    } else {
        for (size_t i = tid; i < nwords; i += stride) p[i] = value;
    }
}
static cudaError_t dev_fill(cudaStream_t s, void* p, int byte_value, size_t bytes) {
    const uint32_t b = (uint32_t) (byte_value & 0xFF), v = b | (b << 8) | (b << 16) | (b << 24);
    const size_t nwords = bytes >> 2;
    if (nwords == 0) return cudaSuccess;
    const unsigned blocks = (unsigned) std::min<size_t>((nwords / 4 + 255) / 256 + 1, 148 * 8);
    k_fill32<<<blocks, 256, 0, s>>>(reinterpret_cast<uint32_t*>(p), v, nwords);
    return cudaGetLastError();
}

static_assert(sizeof(SdmPoint) == 12 && alignof(SdmPoint) == 4, "Point layout (bindings.h:43-47)");
static_assert(sizeof(SdmVoxelField) == 32 && offsetof(SdmVoxelField, voxels) == 16 && offsetof(SdmVoxelField, voxel_count) == 24,
              "VoxelField layout (bindings.h:51-55)");
static_assert(sizeof(SdmVertex) == 24, "Vertex layout (bindings.h:57-60)");
static_assert(sizeof(SdmTriangle) == 72, "Triangle layout (bindings.h:62-64)");
static_assert(sizeof(SdmPrimitive) == 40, "SdmPrimitive layout");

using namespace sdm;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CK(expr)                                                                                                 \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess)                                                                                   \
            return fail(SDM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
    } while (0)

template <class T> struct DevBuf {   // grow-only, like DynamicCudaSlice::get_or_alloc_sync (src/cuda/mod.rs:15-22)
    T* p = nullptr;
    size_t n = 0;
    cudaError_t reserve(size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        else { p = nullptr; (void) cudaGetLastError(); }   // do not leave the allocation failure behind as a sticky "last error"
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
// temporary device buffer of one call: freed on every return path
template <class T> struct TempBuf {
    T* p = nullptr;
    cudaError_t alloc(size_t count) {
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; (void) cudaGetLastError(); }
        return e;
    }
    ~TempBuf() { if (p) cudaFree(p); }
    TempBuf() = default;
    TempBuf(const TempBuf&) = delete;
    TempBuf& operator=(const TempBuf&) = delete;
};

uint32_t grown(uint32_t cap) { return (uint32_t) std::min<uint64_t>((uint64_t) cap * 2, 1ull << 30); }

uint32_t pow2_at_least(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t) std::min<uint64_t>(p, 1ull << 31);
}

// ---- scene compilation ---------------------------------------------------------------------------
// Host float arithmetic here is compiled with -ffp-contract=off: each operation is one IEEE binary32
// operation, the same the reference performs per call on the device (signed_distance.cu:78-79, 94-104).
struct H3 { float x, y, z; };
inline float& at(H3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
inline float at(const H3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

DevPrim make_capsule(H3 b0, H3 b1, float lw, uint32_t fold, float k) {
    DevPrim d;
    memset(&d, 0, sizeof(d));
    const H3 ab { b1.x - b0.x, b1.y - b0.y, b1.z - b0.z };
    const float len = sqrtf(ab.x * ab.x + ab.y * ab.y + ab.z * ab.z);   // length(b1 - b0)
    const H3 bd { ab.x / len, ab.y / len, ab.z / len };                 // (b1 - b0) / len
    d.v0[0] = b0.x; d.v0[1] = b0.y; d.v0[2] = b0.z;
    d.v1[0] = bd.x; d.v1[1] = bd.y; d.v1[2] = bd.z;
    d.v2[0] = b0.x + len * bd.x; d.v2[1] = b0.y + len * bd.y; d.v2[2] = b0.z + len * bd.z;   // bl + len * bd
    d.s0 = lw; d.s1 = len; d.k = k;
    d.kind = SDM_PRIM_CAPSULE; d.fold = fold;
    return d;
}

int compile_scene(const SdmPrimitive* prims, uint32_t count, std::vector<uint4>& blob, bool* has_mandelbulb, float* reach) {
    std::vector<DevPrim> dp;
    *has_mandelbulb = false;
    for (uint32_t i = 0; i < count; i++) {
        const SdmPrimitive& q = prims[i];
        if (q.fold != SDM_FOLD_MIN && q.fold != SDM_FOLD_SMOOTH_MIN) return fail(SDM_ERR_INVALID, "unknown fold op");
        if (q.fold == SDM_FOLD_SMOOTH_MIN && !(q.k > 0.0f)) return fail(SDM_ERR_INVALID, "smooth_min needs k > 0");
        const H3 a { q.a[0], q.a[1], q.a[2] }, b { q.b[0], q.b[1], q.b[2] };
        DevPrim d;
        memset(&d, 0, sizeof(d));
        d.kind = q.kind; d.fold = q.fold; d.k = q.fold == SDM_FOLD_SMOOTH_MIN ? q.k : 0.0f;   // k is only read by smooth folds
        switch (q.kind) {
            case SDM_PRIM_SPHERE:
                d.v0[0] = a.x; d.v0[1] = a.y; d.v0[2] = a.z; d.s0 = q.radius;
                dp.push_back(d);
                break;
            case SDM_PRIM_BOX:
                d.v0[0] = a.x; d.v0[1] = a.y; d.v0[2] = a.z;
                d.v1[0] = b.x / 2.0f; d.v1[1] = b.y / 2.0f; d.v1[2] = b.z / 2.0f;   // bs / 2.0f (signed_distance.cu:87)
                dp.push_back(d);
                break;
            case SDM_PRIM_CAPSULE:
                dp.push_back(make_capsule(a, b, q.radius, q.fold, q.fold == SDM_FOLD_SMOOTH_MIN ? q.k : 0.0f));
                break;
            case SDM_PRIM_BOX_SKELETON: {
                // The skeleton's own fold is min from FLT_MAX (signed_distance.cu:95,109).  Flattening its 12 edges
                // into the scene fold is exact when the scene fold at this point is min too, or when the skeleton
                // is primitive 0 (fold(FLT_MAX, d) = d for both folds).
                if (q.fold != SDM_FOLD_MIN && i != 0)
                    return fail(SDM_ERR_INVALID, "BOX_SKELETON must use fold=min unless it is primitive 0");
                const H3 bs = b;
                const H3 bpl { a.x - bs.x / 2.0f, a.y - bs.y / 2.0f, a.z - bs.z / 2.0f };   // bp - bs / 2.0f
                for (int dir = 0; dir < 3; dir++)
                    for (int c0 = 0; c0 < 2; c0++)
                        for (int c1 = 0; c1 < 2; c1++) {
                            H3 m0 = bpl;
                            at(m0, (dir + 1) % 3) += c0 ? at(bs, (dir + 1) % 2) : 0.0f;   // sic: % 2 (signed_distance.cu:101)
                            at(m0, (dir + 2) % 3) += c1 ? at(bs, (dir + 2) % 3) : 0.0f;
                            H3 m1 = m0;
                            at(m1, dir) += at(bs, dir);
                            dp.push_back(make_capsule(m0, m1, q.radius, SDM_FOLD_MIN, 0.0f));
                        }
            } break;
            case SDM_PRIM_MANDELBULB:
                d.s0 = q.radius;
                *has_mandelbulb = true;
                dp.push_back(d);
                break;
            default:
                return fail(SDM_ERR_INVALID, "unknown primitive kind");
        }
    }
    std::vector<DevRun> runs;
    for (uint32_t i = 0; i < dp.size();) {
        uint32_t j = i + 1;
        while (j < dp.size() && dp[j].kind == dp[i].kind && dp[j].fold == dp[i].fold && j - i < 0xFFFFFFu) j++;
        uint32_t flags = 0;
        if (dp[i].kind == SDM_PRIM_CAPSULE && dp[i].fold == SDM_FOLD_MIN) {
            bool same = true;
            for (uint32_t q = i; q < j; q++) same = same && (memcmp(&dp[q].s0, &dp[i].s0, 4) == 0);
            if (same) flags |= RUN_SHARED_RADIUS_MIN;
        }
        runs.push_back(DevRun { dp[i].kind, dp[i].fold, i, (j - i) | (flags << 24) });
        i = j;
    }
    const size_t bytes = 16 + runs.size() * sizeof(DevRun) + dp.size() * sizeof(DevPrim);
    blob.assign(bytes / 16, make_uint4(0, 0, 0, 0));
    float kmax = 0.0f;
    for (const DevPrim& q : dp) if (q.fold == SDM_FOLD_SMOOTH_MIN) kmax = std::max(kmax, q.k);
    // bound on |d_i(p)| - |p|_1 over the table (k_orient<true>'s rounding bound): |centre| + extent, plus what a smooth-min step can subtract
    double far = 0.0;
    for (const DevPrim& q : dp) {
        const double c = std::fabs((double) q.v0[0]) + std::fabs((double) q.v0[1]) + std::fabs((double) q.v0[2]);
        double ext = 0.0;
        if (q.kind == SDM_PRIM_SPHERE) ext = std::fabs((double) q.s0);
        else if (q.kind == SDM_PRIM_CAPSULE) ext = std::fabs((double) q.s0) + std::fabs((double) q.s1);
        else if (q.kind == SDM_PRIM_BOX) ext = std::fabs((double) q.v1[0]) + std::fabs((double) q.v1[1]) + std::fabs((double) q.v1[2]);
        far = std::max(far, c + ext);
    }
    *reach = (float) (far * 1.001 + (double) kmax + 1.0);   // NaN / inf parameters: the six-sample test then decides nothing
    SceneHeader hdr { (uint32_t) dp.size(), (uint32_t) runs.size(), (uint32_t) bytes, kmax };
    memcpy(blob.data(), &hdr, 16);
    if (!runs.empty()) memcpy(blob.data() + 1, runs.data(), runs.size() * sizeof(DevRun));
    if (!dp.empty()) memcpy(blob.data() + 1 + runs.size(), dp.data(), dp.size() * sizeof(DevPrim));
    return SDM_OK;
}

}  // namespace

struct SdmHandle {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // scene
    DevBuf<uint4> scene;
    uint32_t scene_bytes = 0;
    uint32_t scene_nprims = 0;          // compiled primitives (skeletons expanded)
    bool mask_capable = false;          // large scene of 1-Lipschitz primitives: per-cell primitive masks are used
    bool lipschitz = false;             // no Mandelbulb estimator in the table
    float scene_reach = 0.0f;           // max over the table of |centre|_1 + extent (+ kmax + 1), see k_orient<true>
    DevBuf<uint32_t> masks_fine, masks_coarse;
    DevBuf<uint8_t> cell_maybe;         // per fine cell: 0 = provably no zero crossing inside (k_build_masks)
    MaskGrid grid {};                   // grid.enabled == 0 until ensure_masks has built it
    float grid_bb = 0.0f;
    uint32_t grid_init = 0;
    uint32_t masks_built = 0;           // statistics: number of mask builds

    // field (ping-pong lists) and its host-known description
    DevBuf<float> vox[2];
    int cur = 0;
    int level = 0;              // index into DevState::level_count of the current list
    float voxel_size[3] = { 0, 0, 0 };
    bool have_field = false;
    uint32_t cap_vox = 0;

    // mesh intermediates
    uint32_t cap_tris = 0, cap_uniq = 0, table_entries = 0;
    DevBuf<uint8_t> cases;
    DevBuf<uint32_t> m27;              // per parent: the 27 lattice signs of the last k_refine
    DevBuf<uint32_t> entry_uid;        // per vertex-table entry: the id of its vertex
    // Inherited primitive lists (culled scenes): vl[a] = one record per voxel of the level refined last (its own need-list,
    // k_refine), vl[a ^ 1] = the records of the level before; vparent[b] = record index (= parent) of every voxel of the
    // current level.  lists_level = level whose voxels have a valid vparent (-1: none, the kernels use the cell masks).
    DevBuf<uint4> vl[2];
    DevBuf<uint32_t> vparent[2];
    int vl_cur = 0, vp_cur = 0, lists_level = -1;
    float list_delta = 0.0f;           // inflation of the boxes the current records are proven on (= slack of the mesh stage)
    float delta_override = 0.0f;       // > 0: sdm_remesh knows the final voxel size and uses one inflation for all levels
    DevBuf<uint32_t> urec, tri_rec;    // per vertex / per raw triangle: its list record
    DevBuf<uint32_t> uesc;             // bitmap: vertices that ended outside their record's region
    EdgeLattice lattice {};            // integer lattice of the current field (k_edges' fast keys); enabled = 0: generic keys
    bool lattice_ok = true;            // cleared when a mesh stage reported ERR_LATTICE for the current field
    float field_bb = 0.0f, lat_fail_bb = -1.0f;   // domain of the current field / the one whose grid turned out not to be a lattice
    uint32_t field_init = 0, lat_fail_init = 0;
    DevBuf<uint32_t> vidx;             // per vertex: its index in the welded output
    DevBuf<uint32_t> tri_off, slot_ref, tri_uid, first_slot, wref, first_bits, first_prefix, tri_valid_bits, tri_prefix;
    DevBuf<uint32_t> orient_pending;   // triangles the six-sample orientation test left open (one slot per raw triangle)
    DevBuf<float> ustart, upos, unrm;
    // Mesh outputs are double-buffered: while one mesh is being copied to the host on copy_stream (sdm_mesh_download_async)
    // the next remesh writes the other set.
    DevBuf<uint32_t> out_idx[2];
    DevBuf<float> out_pos[2], out_nrm[2];
    int out_sel = 0;                     // set the NEXT weld writes
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_mesh_done[2] = { nullptr, nullptr }, ev_copy_done[2] = { nullptr, nullptr };
    DevBuf<uint4> table1, table2;
    DevBuf<uint64_t> tiles;
    DevBuf<Straggler> stragglers;
    uint32_t cap_stragglers = 0;
    uint32_t cases_epoch = 0;           // epoch passed to that k_refine (DevState::cases_from_refine)
    int cases_for_level = -1;           // level whose case indices the last k_refine wrote (-1: none)
    uint32_t table1_entries = 0;        // entries of table1 actually used (adaptive: sized from the previous mesh)
    DevBuf<DevState> state;
    DevState* host_state = nullptr;   // pinned
    DevBuf<uchar4> render_target;     // sdm_render's RGBA8 image (grow-only, like render_texture_buffer, src/cuda/mod.rs:33)
    DevBuf<uint32_t> shard_range;     // {lo, hi, n} of the last k_take_shard
    uint32_t* host_range = nullptr;   // pinned
    uint32_t shard_index = 0;
    int welded_set = 0, welded_vset = 0;   // output sets holding the local weld's indices / vertices
    uint32_t welded_v = 0, welded_t = 0;   // rows of the local weld (after sdm_shard_apply_remap: kept vertices)
    ShardOffsets res_offsets {};      // vertex / duplicate-pair offsets of the last sdm_shard_resolve
    // peer exchange (sdm_peer_*): rank 0's control block / row slots / pair lists / second output set - own or IPC-mapped
    DevBuf<unsigned char> peer_block;
    PeerCtl* ctl = nullptr;
    uint4* peer_rows = nullptr;
    uint2* peer_pairs = nullptr;
    float *root_pos = nullptr, *root_nrm = nullptr;
    uint32_t* root_idx = nullptr;
    uint32_t peer_rank = 0, peer_world = 0, peer_cap_rows = 0, peer_cap_v = 0, peer_cap_t = 0;
    void* ipc_mapped[4] = { nullptr, nullptr, nullptr, nullptr };
    DevBuf<float> peer_frac;          // cumulative weight fractions of the shards, two generations of SDM_PEER_MAX + 1 (k_peer_rebalance)
    DevBuf<unsigned long long> peer_t0;
    const float* shard_frac = nullptr;   // fractions the next enqueue_take_shard uses (null: equal shares)
    uint32_t peer_last_epoch = 0;
    bool peer_rebalance = true;       // SDM_NO_REBALANCE switches the measured load balancing off
    DevBuf<PeerLocal> peer_local;
    PeerLocal* host_peer_local = nullptr;   // pinned
    int peer_deliver = 0;
    DevBuf<uint32_t> shard_scratch;   // 128 words: x range, counters, per-shard duplicate counts / cursors / offsets
    uint32_t* host_scratch = nullptr; // pinned mirror
    uint32_t own_tris = 0, own_uniq = 0;   // the local shard's counts (sdm_shard_remesh)
    uint32_t epoch = 1;
    bool mesh_valid = false;

    // persistent grid sizes
    int g_refine = 0, g_classify = 0, g_project = 0, g_tail = 0, g_normals = 0, g_orient = 0, g_orient_quick = 0, g_orient_pending = 0, g_light = 0, g_edges = 0;
    uint32_t proj_chunk = 256;          // vertex chunk per warp in k_project (SDM_PROJ_CHUNK overrides)
    float slack_factor = 1.0f;         // inflation of the list regions in units of the child voxel size (SDM_SLACK overrides)
    bool use_lists = true, use_lattice = true, use_masks = true, use_quick_orient = true;   // SDM_NO_LISTS / SDM_NO_LATTICE / SDM_NO_MASKS: developer switches (A/B measurements, tests)

    SdmStats stats {};

    // optional per-kernel timing (sdm_set_profiling): one event after every enqueued kernel / clear
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<std::string> prof_names;
    size_t prof_used = 0;
    std::vector<float> prof_ms;
};

namespace {

// dynamic shared memory: culled scenes need one culling slot per warp (the table is read through L1); small scenes
// stage the table.  k_build_masks always stages it (staged = true).
size_t smem_for(const SdmHandle* h, int threads = 256, bool staged = false) {
    size_t b;
    if (h->mask_capable && !staged) b = (size_t) (threads / 32) * cull_smem_per_warp((h->scene_nprims + 31) / 32);
    else b = (size_t) h->scene_bytes;
    return (b + 15) & ~(size_t) 15;
}

// NVTX ranges per remesh, level and stage (the reference has them on its render path only,
// src/cuda/mod.rs:354-408).  Header-only NVTX3: a no-op unless a tool (nsys / ncu --nvtx) is attached.
struct NvtxRange {
    NvtxRange(const SdmHandle*, const char* name, int index = -1) {
        if (index < 0) { nvtxRangePushA(name); return; }
        char buf[64];
        snprintf(buf, sizeof buf, "%s %d", name, index);
        nvtxRangePushA(buf);
    }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// profiling: mark(h, name) closes the interval that started at the previous mark
void prof_begin(SdmHandle* h) {
    if (!h->profiling) return;
    h->prof_used = 0;
    h->prof_names.clear();
}
void mark(SdmHandle* h, const char* name) {
    if (!h->profiling) return;
    if (h->prof_used == h->prof_events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->prof_events.push_back(e);
    }
    cudaEventRecord(h->prof_events[h->prof_used++], h->stream);
    h->prof_names.push_back(name);
}
void prof_end(SdmHandle* h) {   // after the stream has been synchronised
    h->prof_ms.clear();
    if (!h->profiling) return;
    for (size_t i = 1; i < h->prof_used; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->prof_events[i - 1], h->prof_events[i]);
        h->prof_ms.push_back(ms);
    }
}

int configure_kernels(SdmHandle* h) {
    if (!h->mask_capable && smem_for(h, 256, true) > 200 * 1024)
        return fail(SDM_ERR_INVALID, "an un-culled scene table must fit in shared memory (max ~3000 primitives)");
    if (h->scene_nprims > 65535) return fail(SDM_ERR_INVALID, "too many primitives (tile lists hold 16-bit indices)");
    struct K { const void* f; int threads; int* grid; };
    const K ks[] = {
        { (const void*) k_refine, 256, &h->g_refine },   { (const void*) k_cases, 256, &h->g_classify },
        { (const void*) k_project, 128, &h->g_project }, { (const void*) k_vertex_normals, 128, &h->g_normals },
        { (const void*) k_orient<false>, 128, &h->g_orient },   { (const void*) k_project_tail, 128, &h->g_tail },
        { (const void*) k_orient<true>, 128, &h->g_orient_quick }, { (const void*) k_orient_pending, 128, &h->g_orient_pending },
    };
    for (const K& k : ks) {
        const size_t smem = smem_for(h, k.threads);
        CK(cudaFuncSetAttribute(k.f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 1024)));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k.f, k.threads, smem));
        if (per_sm < 1) return fail(SDM_ERR_CUDA, "kernel does not fit on an SM");
        *k.grid = per_sm * h->num_sms;
    }
    for (const void* f : { (const void*) k_eval_sdf, (const void*) k_eval_normal, (const void*) k_eval_project, (const void*) k_render })
        CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem_for(h, 128), 1024)));
    h->g_light = h->num_sms * 8;
    {   // static-shared-memory kernels of the classification stage
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_edges<true>, 256, 0));
        h->g_edges = std::max(per_sm, 1) * h->num_sms;
    }
    return SDM_OK;
}

// Per-cell primitive masks over the cube [-bb/2, bb/2]^3 (two levels: G/4 coarse cells prune for the G fine cells).
int ensure_masks(SdmHandle* h, float bb_size, uint32_t init_factor) {
    if (!h->mask_capable) { h->grid.enabled = 0; return SDM_OK; }
    // The culling tests prove their drops with an absolute margin of 1e-4 against float rounding of distances (~3e-7 |d|):
    // sound for the domains this path is used on (the reference's cube is 5 wide), not for arbitrarily large ones.
    if (!(bb_size <= 32.0f)) return fail(SDM_ERR_INVALID, "scenes of more than 24 primitives need bb_size <= 32 (culling margins are absolute)");
    const uint32_t W = (h->scene_nprims + 31) / 32;
    uint32_t G = 16;
    while (G < init_factor && G < 128) G <<= 1;
    if (W > 32 && G > 64) G = 64;
    if (h->grid.enabled && h->grid.G == G && h->grid_bb == bb_size) return SDM_OK;
    const uint32_t Gc = G / 4;
    CK(h->masks_fine.reserve((size_t) G * G * G * W));
    CK(h->masks_coarse.reserve((size_t) Gc * Gc * Gc * W));
    MaskGrid fine {};
    fine.masks = h->masks_fine.p; fine.G = G; fine.W = W;
    fine.ox = fine.oy = fine.oz = -bb_size / 2.0f;
    fine.cell = bb_size / (float) G; fine.inv_cell = (float) G / bb_size; fine.enabled = 1;
    MaskGrid coarse = fine;
    coarse.masks = h->masks_coarse.p; coarse.G = Gc; coarse.cell = bb_size / (float) Gc; coarse.inv_cell = (float) Gc / bb_size;
    // radius = circumsphere of the cell cube (x1.0001) + the empirical_normal stencil reach (2e-3, signed_distance.cu:179)
    //          + slop for the inward-nudged box probes and the domain-face tolerance (2e-3 cell) + 1e-4
    auto rho = [](float cell) { return cell * 0.8660254f * 1.0001f + 0.0021f + 2e-3f * cell + 1e-4f; };
    CK(h->cell_maybe.reserve((size_t) G * G * G));
    fine.maybe = nullptr; coarse.maybe = nullptr;
    k_build_masks<<<h->num_sms * 8, 256, 0, h->stream>>>(h->scene.p, h->masks_coarse.p, coarse, nullptr, 0, rho(coarse.cell), nullptr);
    if (W <= 32)
        k_build_masks_fine<<<h->num_sms * 4, 256, 0, h->stream>>>(h->scene.p, h->masks_fine.p, fine, h->masks_coarse.p, Gc, rho(fine.cell), h->cell_maybe.p);
    else
        k_build_masks<<<h->num_sms * 8, 256, 0, h->stream>>>(h->scene.p, h->masks_fine.p, fine, h->masks_coarse.p, Gc, rho(fine.cell), h->cell_maybe.p);
    fine.maybe = h->cell_maybe.p;
    mark(h, "k_build_masks_x2");
    h->stats.kernel_launches += 2;
    CK(cudaGetLastError());
    h->grid = fine;
    h->grid_bb = bb_size;
    h->grid_init = init_factor;
    h->masks_built++;
    return SDM_OK;
}

// All buffers that scale with the voxel capacity.  The capacities in the handle are committed only after every allocation
// has succeeded; if one fails (DevBuf::reserve has already freed the old buffer) they are zeroed, so that the next call
// re-allocates instead of launching kernels bounded by capacities that no buffer has.
int reserve_all(SdmHandle* h, uint32_t cap_vox, uint32_t cap_tris, uint32_t cap_uniq, uint32_t table_entries) {
    for (int i = 0; i < 2; i++) CK(h->vox[i].reserve((size_t) cap_vox * 3));
    CK(h->cases.reserve((size_t) cap_vox + 4));
    for (int i = 0; i < 2; i++) { CK(h->vl[i].reserve((size_t) cap_vox * 2)); CK(h->vparent[i].reserve(cap_vox)); }
    CK(h->tri_rec.reserve(cap_tris));
    CK(h->urec.reserve(cap_uniq));
    CK(h->uesc.reserve((size_t) cap_uniq / 32 + 32));
    CK(h->m27.reserve(cap_vox));
    CK(h->tri_off.reserve(cap_vox));
    CK(h->slot_ref.reserve((size_t) cap_tris * 3));
    CK(h->tri_uid.reserve((size_t) cap_tris * 3));
    for (int b = 0; b < 2; b++) CK(h->out_idx[b].reserve((size_t) cap_tris * 3));
    CK(h->first_bits.reserve(((size_t) cap_tris * 3 + 31) / 32 + 32));
    CK(h->first_prefix.reserve(((size_t) cap_tris * 3 + 31) / 32 + 32));
    CK(h->tri_valid_bits.reserve(((size_t) cap_tris + 31) / 32 + 32));
    CK(h->orient_pending.reserve((size_t) cap_tris + 32));
    CK(h->tri_prefix.reserve(((size_t) cap_tris + 31) / 32 + 32));
    CK(h->first_slot.reserve(cap_uniq));
    CK(h->wref.reserve(cap_uniq));
    CK(h->vidx.reserve(cap_uniq));
    CK(h->ustart.reserve((size_t) cap_uniq * 3));
    CK(h->upos.reserve((size_t) cap_uniq * 3));
    CK(h->unrm.reserve((size_t) cap_uniq * 3));
    for (int b = 0; b < 2; b++) { CK(h->out_pos[b].reserve((size_t) cap_uniq * 3)); CK(h->out_nrm[b].reserve((size_t) cap_uniq * 3)); }
    CK(h->table1.reserve(table_entries));
    CK(h->entry_uid.reserve(table_entries));
    CK(h->table2.reserve(table_entries));
    const size_t max_tiles = std::max<size_t>(((size_t) cap_tris * 3 / 32 + 31) / 32, (size_t) cap_vox / 32) + 64;
    const size_t old_tiles = h->tiles.n;
    CK(h->tiles.reserve(max_tiles));
    if (h->tiles.n != old_tiles) CK(dev_fill(h->stream, h->tiles.p, 0, h->tiles.n * sizeof(uint64_t)));
    CK(h->stragglers.reserve(cap_uniq / 8 + 4096));
    return SDM_OK;
}

int ensure_capacity(SdmHandle* h, uint32_t cap_vox) {
    if (cap_vox <= h->cap_vox) return SDM_OK;
    if (h->copy_stream) CK(cudaStreamSynchronize(h->copy_stream));   // no download may be reading a buffer that is about to move
    // copy-preserving growth of the current list is not needed: callers re-create the field after growing
    const uint32_t cap_tris = (uint32_t) std::min<uint64_t>((uint64_t) cap_vox * 3, 0x3FFFFFFFull);   // < 2^32 / 3 slots
    const uint32_t cap_uniq = (uint32_t) std::min<uint64_t>((uint64_t) cap_vox * 2, 0x7FFFFFFFull);
    const uint32_t table_entries = pow2_at_least((uint64_t) cap_uniq * 2);
    h->have_field = false;
    h->mesh_valid = false;
    const int rc = reserve_all(h, cap_vox, cap_tris, cap_uniq, table_entries);
    if (rc) {
        h->cap_vox = h->cap_tris = h->cap_uniq = h->table_entries = h->table1_entries = h->cap_stragglers = 0;
        return rc;
    }
    h->cap_vox = cap_vox; h->cap_tris = cap_tris; h->cap_uniq = cap_uniq; h->table_entries = table_entries;
    h->cap_stragglers = cap_uniq / 8 + 4096;
    h->table1_entries = h->table_entries;
    return SDM_OK;
}

// capacity (a doubling of the current one) that holds n0 level-0 voxels; the lists are limited to 2^30 voxels
int capacity_for(SdmHandle* h, uint64_t n0, uint32_t* want) {
    if (n0 > (1ull << 30)) return fail(SDM_ERR_CAPACITY, "more than 2^30 voxels in one list");
    uint32_t w = std::max(h->cap_vox, 1u << 21);
    while (w < n0) w = grown(w);
    *want = w;
    return SDM_OK;
}

uint32_t next_epoch(SdmHandle* h) {
    h->epoch++;
    if (h->epoch >= (1u << 30)) {   // wrap: make every stale descriptor invalid again
        dev_fill(h->stream, h->tiles.p, 0, h->tiles.n * sizeof(uint64_t));
        h->epoch = 1;
    }
    return h->epoch;
}

int reset_state(SdmHandle* h) {
    CK(dev_fill(h->stream, h->state.p, 0, sizeof(DevState)));
    return SDM_OK;
}

// culled scenes must never be launched without a grid (their kernels do not stage the table): rebuild over the last
// domain, or the reference's default cube (bindings.h:9-10); points outside simply use the full list
int ensure_masks_any(SdmHandle* h) {
    if (!h->mask_capable || h->grid.enabled) return SDM_OK;
    return ensure_masks(h, h->grid_bb > 0.0f ? h->grid_bb : SDM_MESH_GENERATION_BB_SIZE, h->grid_init ? h->grid_init : 64);
}

int enqueue_init_field(SdmHandle* h, const SdmParams& p) {
    int mrc = ensure_masks(h, p.bb_size, p.init_factor);
    if (mrc) return mrc;
    const float size = p.bb_size / (float) p.init_factor;   // src/cuda/mod.rs:106
    k_init_field<<<h->g_light, 256, 0, h->stream>>>(h->vox[0].p, h->state.p, p.bb_size, p.init_factor, size, h->cap_vox);
    mark(h, "k_init_field");
    h->stats.kernel_launches++;
    h->cur = 0; h->level = 0;
    h->cases_for_level = -1;
    h->voxel_size[0] = h->voxel_size[1] = h->voxel_size[2] = size;
    h->have_field = true;
    h->mesh_valid = false;
    h->lists_level = -1;
    // the field's lattice: origin = min corner of the domain; the unit (half a voxel) follows the voxel size at mesh time
    h->lattice.ox = h->lattice.oy = h->lattice.oz = 0.0f - p.bb_size / 2.0f;
    h->lattice.enabled = 1;
    h->field_bb = p.bb_size; h->field_init = p.init_factor;
    h->lattice_ok = !(p.bb_size == h->lat_fail_bb && p.init_factor == h->lat_fail_init);
    return SDM_OK;
}

// with_cases: this is the refinement right before a mesh stage - let it also write the children's case indices
int enqueue_refine(SdmHandle* h, bool with_cases = false) {
    if (h->level >= 15) return fail(SDM_ERR_INVALID, "too many refinement levels (max 15)");
    int mrc = ensure_masks_any(h);
    if (mrc) return mrc;
    const float ox = h->voxel_size[0] / 2.0f, oy = h->voxel_size[1] / 2.0f, oz = h->voxel_size[2] / 2.0f;   // :20
    // Inherited lists (culled scenes): this level's parents get their own records (vl[vl_cur ^ 1]), proven on their boxes inflated
    // by delta; candidates come from the records they inherited (vl[vl_cur] through vparent[vp_cur]) when there are any.
    // Level 0 is dense: a tile's 32 parents are 32 different cells, and inflating them would pull the rows of ~300 cells into the
    // tile's candidates.  Records start at level 1 (whose parents take their candidates from the cell masks once more).
    const bool lists = h->mask_capable && h->grid.enabled && h->use_lists && h->grid.W <= 32 && h->level >= 1;
    const bool inherit = lists && h->lists_level == h->level;
    const float delta = lists ? (h->delta_override > 0.0f ? h->delta_override : h->slack_factor * std::max(ox, std::max(oy, oz))) : 0.0f;
    NvtxRange nv(h, "refine level", h->level);
    k_refine<<<h->g_refine, 256, smem_for(h, 256), h->stream>>>(h->scene.p, h->vox[h->cur].p, h->state.p, h->level, ox, oy, oz, h->grid, h->m27.p,
                                                                with_cases ? (h->cases_epoch = next_epoch(h)) : 0u, h->level == 0 && h->grid.enabled ? 1 : 0,
                                                                inherit ? h->vl[h->vl_cur].p : nullptr, inherit ? h->vparent[h->vp_cur].p : nullptr,
                                                                lists ? h->vl[h->vl_cur ^ 1].p : nullptr, delta);
    mark(h, "k_refine");
    k_refine_emit<<<h->g_light, 256, 0, h->stream>>>(h->vox[h->cur].p, h->vox[h->cur ^ 1].p, h->state.p, h->level, next_epoch(h), h->tiles.p, h->cap_vox,
                                                     ox, oy, oz, h->m27.p, with_cases ? h->cases.p : nullptr, lists ? h->vparent[h->vp_cur ^ 1].p : nullptr);
    if (lists) { h->vl_cur ^= 1; h->vp_cur ^= 1; h->lists_level = h->level + 1; h->list_delta = delta; }
    else h->lists_level = -1;
    h->cases_for_level = with_cases ? h->level + 1 : -1;
    mark(h, "k_refine_emit");
    h->stats.kernel_launches += 2;
    h->cur ^= 1; h->level++;
    h->voxel_size[0] = ox; h->voxel_size[1] = oy; h->voxel_size[2] = oz;
    h->mesh_valid = false;
    return SDM_OK;
}

// clears sized on the device from n_uniq / n_tris_raw (which must already be in DevState)
// this rank's contiguous part of the current list -> the other ping-pong buffer.  At level 0 of a culled scene the parts are
// balanced by the cells' surface flags (no level needs to be refined redundantly on every rank); otherwise equal parts.
int enqueue_take_shard(SdmHandle* h, uint32_t shard_index, uint32_t shard_count) {
    const bool vp = h->lists_level == h->level;   // the shard's voxels keep their record indices
    const uint32_t* bounds = nullptr;
    if (h->level == 0 && h->grid.enabled && h->grid.maybe && h->grid.G == h->field_init && shard_count > 1) {
        k_shard_bounds_by_flags<<<1, 1024, 0, h->stream>>>(h->grid.maybe, h->grid.G * h->grid.G * h->grid.G, shard_index, shard_count, h->shard_range.p + 4,
                                                            h->shard_frac);
        bounds = h->shard_range.p + 4;
        h->stats.kernel_launches++;
    }
    k_take_shard<<<h->g_light, 256, 0, h->stream>>>(h->vox[h->cur].p, h->vox[h->cur ^ 1].p, h->state.p, h->level, shard_index, shard_count, h->shard_range.p,
                                                     vp ? h->vparent[h->vp_cur].p : nullptr, vp ? h->vparent[h->vp_cur ^ 1].p : nullptr, bounds);
    k_set_level_count<<<1, 1, 0, h->stream>>>(h->state.p, h->level, h->shard_range.p);
    mark(h, "k_take_shard");
    h->stats.kernel_launches += 2;
    h->cur ^= 1;
    if (vp) h->vp_cur ^= 1;
    return SDM_OK;
}

int enqueue_weld_clears(SdmHandle* h, bool clear_first_slot) {
    k_clear_weld_state<<<h->g_light, 256, 0, h->stream>>>(h->state.p, h->first_slot.p, h->first_bits.p, h->table2.p, h->table_entries,
                                                          h->cap_uniq, clear_first_slot ? 1 : 0, clear_first_slot ? h->uesc.p : nullptr);
    h->stats.kernel_launches++;
    return SDM_OK;
}

// classify+edges -> project (+tail) -> normals -> orient: everything that needs only this handle's voxels
int enqueue_mesh_local(SdmHandle* h, bool fuse_weld_keys) {
    int mrc = ensure_masks_any(h);
    if (mrc) return mrc;
    const float sx = h->voxel_size[0], sy = h->voxel_size[1], sz = h->voxel_size[2];
    const float* vox = h->vox[h->cur].p;
    const size_t smem = smem_for(h, 256), smem128 = smem_for(h, 128);
    cudaStream_t s = h->stream;
    // the mesh stage may be re-run on the same field: reset the mesh-stage counters and tickets only (not error_flags)
    k_reset_mesh_state<<<1, 32, 0, s>>>(h->state.p);
    // Vertex de-duplication table: sized from the previous mesh of this handle (4x its vertex count), full size at first; an
    // overflow is detected (ERR_HASH_FULL) and retried with the full table.  Fast path: 8-byte lattice keys (0 = empty).
    const bool lat = h->use_lattice && h->lattice.enabled && h->lattice_ok;
    EdgeLattice L = h->lattice;
    L.hx = sx / 2.0f; L.hy = sy / 2.0f; L.hz = sz / 2.0f;
    L.ihx = 1.0f / L.hx; L.ihy = 1.0f / L.hy; L.ihz = 1.0f / L.hz;
    CK(dev_fill(s, h->table1.p, lat ? 0x00 : 0xFF, (size_t) h->table1_entries * (lat ? 8 : 16)));
    mark(h, "clears");
    const bool lists = h->mask_capable && h->use_lists && h->lists_level == h->level && h->grid.W <= 32;
    const uint4* vl = lists ? h->vl[h->vl_cur].p : nullptr;
    const uint32_t* vparent = lists ? h->vparent[h->vp_cur].p : nullptr;
    const float slack2 = h->list_delta * h->list_delta;
    {
        NvtxRange nv(h, "mesh: classify + edge vertices");
        k_cases<<<h->g_classify, 256, smem, s>>>(h->scene.p, vox, h->state.p, h->level, h->cases.p, sx, sy, sz, h->grid, h->cases_for_level == h->level ? h->cases_epoch : 0u);
        k_tri_offsets<<<h->g_light, 256, 0, s>>>(h->state.p, h->level, h->cases.p, h->tri_off.p, next_epoch(h), h->tiles.p, h->cap_tris);
        mark(h, "k_cases+k_tri_offsets");
        if (lat)
            k_edges<true><<<h->g_edges, 256, 0, s>>>(vox, h->state.p, h->level, h->cases.p, h->tri_off.p, h->table1.p, h->table1_entries - 1, h->slot_ref.p, h->tri_rec.p,
                                                     vparent, h->entry_uid.p, h->ustart.p, h->urec.p, L, sx, sy, sz, h->cap_uniq);
        else
            k_edges<false><<<h->g_edges, 256, 0, s>>>(vox, h->state.p, h->level, h->cases.p, h->tri_off.p, h->table1.p, h->table1_entries - 1, h->slot_ref.p, h->tri_rec.p,
                                                      vparent, h->entry_uid.p, h->ustart.p, h->urec.p, L, sx, sy, sz, h->cap_uniq);
        mark(h, "k_edges");
        h->stats.kernel_launches += 3;
    }
    int rc = enqueue_weld_clears(h, true);
    if (rc) return rc;
    mark(h, "k_clear_weld_state");
    {
        NvtxRange nv(h, "mesh: project");
        k_project<<<h->g_project, 128, smem128, s>>>(h->scene.p, h->state.p, h->ustart.p, h->upos.p, h->cap_uniq, h->stragglers.p, h->cap_stragglers, h->grid,
                                                     h->proj_chunk, vl, h->urec.p, h->uesc.p, slack2);
        mark(h, "k_project");
        k_project_tail<<<h->g_tail, 128, smem128, s>>>(h->scene.p, h->state.p, h->upos.p, h->stragglers.p, h->cap_stragglers, h->grid, h->ustart.p,
                                                       vl ? h->uesc.p : nullptr, slack2);
        mark(h, "k_project_tail");
    }
    {
        NvtxRange nv(h, "mesh: vertex normals + weld keys");
        k_vertex_normals<<<h->g_normals, 128, smem128, s>>>(h->scene.p, h->state.p, h->upos.p, h->unrm.p, h->cap_uniq, h->grid,
                                                            fuse_weld_keys ? h->table2.p : nullptr, h->table_entries, h->wref.p, vl, h->urec.p, h->uesc.p);
        mark(h, "k_vertex_normals");
    }
    {
        NvtxRange nv(h, "mesh: orient");
        if (h->lipschitz && h->use_quick_orient) {
            k_orient<true><<<h->g_orient_quick, 128, smem128, s>>>(h->scene.p, h->state.p, h->entry_uid.p, h->slot_ref.p, h->upos.p, h->tri_uid.p, h->first_slot.p,
                                                              h->tri_valid_bits.p, h->grid, vl, h->tri_rec.p, h->uesc.p, h->orient_pending.p, h->scene_reach);
            k_orient_pending<<<h->g_orient_pending, 128, smem128, s>>>(h->scene.p, h->state.p, h->entry_uid.p, h->slot_ref.p, h->upos.p, h->tri_uid.p,
                                                                       h->first_slot.p, h->tri_valid_bits.p, h->grid, vl, h->tri_rec.p, h->uesc.p, h->orient_pending.p);
            h->stats.kernel_launches++;
        } else {
            k_orient<false><<<h->g_orient, 128, smem128, s>>>(h->scene.p, h->state.p, h->entry_uid.p, h->slot_ref.p, h->upos.p, h->tri_uid.p, h->first_slot.p,
                                                               h->tri_valid_bits.p, h->grid, vl, h->tri_rec.p, h->uesc.p, nullptr, 0.0f);
        }
        mark(h, "k_orient");
    }
    h->stats.kernel_launches += 6;
    CK(cudaGetLastError());
    return SDM_OK;
}

// the reference-order weld over (upos, unrm, tri_uid, first_slot, tri_valid_bits) and the counters in DevState
int enqueue_weld(SdmHandle* h, bool keys_inserted) {
    NvtxRange nv(h, "weld + emit");
    cudaStream_t s = h->stream;
    const int b = h->out_sel;
    CK(cudaStreamWaitEvent(s, h->ev_copy_done[b], 0));   // the download of the mesh that used this set two remeshes ago
    if (!keys_inserted) {
        k_weld_keys<<<h->g_light, 256, 0, s>>>(h->state.p, h->upos.p, h->table2.p, h->table_entries, h->wref.p, h->cap_uniq);
        h->stats.kernel_launches++;
    }
    k_weld_min<<<h->g_light, 256, 0, s>>>(h->state.p, h->first_slot.p, h->table2.p, h->wref.p, h->cap_uniq);
    k_weld_mark<<<h->g_light, 256, 0, s>>>(h->state.p, h->first_slot.p, h->table2.p, h->wref.p, h->first_bits.p, h->cap_uniq);
    mark(h, "k_weld_min+mark");
    k_bitscan<<<h->g_light, 256, 0, s>>>(h->state.p, h->first_bits.p, h->first_prefix.p, 0, next_epoch(h), h->tiles.p);
    k_bitscan<<<h->g_light, 256, 0, s>>>(h->state.p, h->tri_valid_bits.p, h->tri_prefix.p, 1, next_epoch(h), h->tiles.p);
    mark(h, "k_bitscan_x2");
    k_emit_vertices<<<h->g_light, 256, 0, s>>>(h->state.p, h->first_slot.p, h->table2.p, h->wref.p, h->first_bits.p, h->first_prefix.p,
                                                h->upos.p, h->unrm.p, h->out_pos[b].p, h->out_nrm[b].p, h->vidx.p, h->cap_uniq);
    mark(h, "k_emit_vertices");
    k_emit_indices<<<h->g_light, 256, 0, s>>>(h->state.p, h->tri_uid.p, h->vidx.p, h->tri_valid_bits.p, h->tri_prefix.p, h->out_idx[b].p);
    mark(h, "k_emit_indices");
    CK(cudaEventRecord(h->ev_mesh_done[b], s));
    h->stats.kernel_launches += 6;
    CK(cudaGetLastError());
    return SDM_OK;
}

int enqueue_mesh(SdmHandle* h) {
    int rc = enqueue_mesh_local(h, true);
    if (rc) return rc;
    return enqueue_weld(h, true);
}

// Copies DevState to the pinned host mirror and waits.  Returns the device-side error flags through *flags.
int fetch_state(SdmHandle* h, uint32_t* flags) {
    CK(cudaMemcpyAsync(h->host_state, h->state.p, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *flags = h->host_state->error_flags;
    return SDM_OK;
}

void fill_stats(SdmHandle* h, bool meshed) {
    const DevState& st = *h->host_state;
    for (int i = 0; i < 16; i++) h->stats.level_counts[i] = st.level_count[i];
    if (meshed) {
        h->stats.unique_vertices = st.n_uniq;
        h->stats.raw_triangles = st.n_tris_raw;
    }
    // analytic evaluation count: 27 per refined parent, 8 per meshed voxel, 13 per Newton iteration,
    // 12 per vertex normal, 12 per triangle (centroid normal)
    uint64_t evals = 0;
    for (int l = 0; l < h->level; l++) evals += 27ull * st.level_count[l];
    if (meshed) evals += 8ull * st.level_count[h->level] + 13ull * st.newton_iters + 12ull * st.n_uniq + 12ull * st.n_tris_raw;
    h->stats.sdf_evals = evals;
    for (int i = 0; i < 6; i++) h->stats.prim_evals[i] = st.prim_evals[i];
    if (meshed) {
        h->stats.escaped_vertices = st.n_escaped; h->stats.list_fallback_tiles = st.list_fallbacks;
        h->stats.stragglers = st.n_stragglers; h->stats.newton_iterations = st.newton_iters;
    }
}

void mesh_view(SdmHandle* h, SdmMesh* m) {
    const int b = h->out_sel;          // the set the weld that just finished wrote
    m->positions = h->out_pos[b].p;
    m->normals = h->out_nrm[b].p;
    m->indices = h->out_idx[b].p;
    h->out_sel ^= 1;
    m->vertex_count = h->host_state->n_verts_out;
    m->triangle_count = h->host_state->n_tris_out;
    m->on_device = 1;
    m->reserved = b;
}

int check_params(const SdmParams& p) {
    if (!(p.bb_size > 0.0f) || p.init_factor == 0 || p.init_factor > 1024 || p.levels > 15)
        return fail(SDM_ERR_INVALID, "bad SdmParams");
    return SDM_OK;
}

// after a successful mesh: size the vertex table of the next mesh from this one
void adapt_table1(SdmHandle* h) {
    const uint32_t want = pow2_at_least(std::max<uint64_t>((uint64_t) h->host_state->n_uniq * 7 / 4, 1u << 16));
    h->table1_entries = std::min(want, h->table_entries);
}
// the adaptive vertex table was too small (and nothing else overflowed): retry with the full table, same capacities
bool only_table1_overflow(SdmHandle* h, uint32_t flags) {
    if (flags == ERR_HASH_FULL && h->table1_entries < h->table_entries) { h->table1_entries = h->table_entries; return true; }
    // a voxel off the integer lattice (non-dyadic grid): same capacities, generic float-bit keys from now on for this field
    if (flags == ERR_LATTICE && h->lattice_ok) { h->lattice_ok = false; h->lat_fail_bb = h->field_bb; h->lat_fail_init = h->field_init; return true; }
    return false;
}

}  // namespace

extern "C" {

const char* sdm_last_error(void) { return g_last_error.c_str(); }
const char* sdm_version(void) { return "sdfmesh-b200 0.1 (sm_100a)"; }

uint32_t sdm_scene_default(SdmPrimitive* out, uint32_t capacity) {
    if (out && capacity >= 2) {
        memset(out, 0, 2 * sizeof(SdmPrimitive));
        // common.cu:222-226: smooth_min(sd_box_skeleton(p, vec3(0), vec3(3, 1, .5), .1), length(p) - 1, .5)
        out[0].kind = SDM_PRIM_BOX_SKELETON; out[0].fold = SDM_FOLD_MIN; out[0].radius = 0.1f;
        out[0].b[0] = 3.0f; out[0].b[1] = 1.0f; out[0].b[2] = 0.5f;
        out[1].kind = SDM_PRIM_SPHERE; out[1].fold = SDM_FOLD_SMOOTH_MIN; out[1].k = 0.5f; out[1].radius = 1.0f;
    }
    return 2;
}

int sdm_set_scene(SdmHandle* h, const SdmPrimitive* prims, uint32_t count) {
    if (!h || (!prims && count)) return fail(SDM_ERR_INVALID, "null argument");
    std::vector<uint4> blob;
    bool mb = false;
    float reach = 0.0f;
    int rc = compile_scene(prims, count, blob, &mb, &reach);
    if (rc) return rc;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));   // the previous table may still be in use
    CK(h->scene.reserve(blob.size()));
    CK(cudaMemcpyAsync(h->scene.p, blob.data(), blob.size() * 16, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // `blob` is a local
    const uint32_t old_bytes = h->scene_bytes, old_nprims = h->scene_nprims;
    const bool old_capable = h->mask_capable;
    h->scene_bytes = (uint32_t) (blob.size() * 16);
    h->scene_nprims = reinterpret_cast<const SceneHeader*>(blob.data())->nprims;
    // culling pays off once the table is long; it needs 1-Lipschitz primitives (not the Mandelbulb estimator)
    h->mask_capable = !mb && h->scene_nprims > 24 && h->use_masks;
    h->lipschitz = !mb;   // sphere / capsule / box folded by min / smooth-min: the six-sample orientation test applies
    h->scene_reach = reach;
    h->grid.enabled = 0;   // masks depend on the scene: rebuilt on the next remesh
    h->mesh_valid = false;
    // shared-memory sizes / persistent grid sizes only depend on these three: an animated scene (same table shape every
    // frame) does not pay for the occupancy queries again
    if (h->g_refine && old_bytes == h->scene_bytes && old_nprims == h->scene_nprims && old_capable == h->mask_capable) return SDM_OK;
    return configure_kernels(h);
}

int sdm_create(int device_ordinal, SdmHandle** out_handle) {
    if (!out_handle) return fail(SDM_ERR_INVALID, "null out_handle");
    *out_handle = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SDM_ERR_NO_DEVICE, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device_ordinal < 0 || device_ordinal >= ndev) return fail(SDM_ERR_INVALID, "device ordinal out of range");
    CK(cudaSetDevice(device_ordinal));
    SdmHandle* h = new SdmHandle();
    h->device = device_ordinal;
    int rc = SDM_OK;
    do {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { rc = fail(SDM_ERR_CUDA, "cudaGetDeviceProperties"); break; }
        h->num_sms = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(SDM_ERR_CUDA, "stream"); break; }
        cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
        cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
        for (int b = 0; b < 2; b++) { cudaEventCreateWithFlags(&h->ev_mesh_done[b], cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_copy_done[b], cudaEventDisableTiming); }
        if (cudaMallocHost(&h->host_state, sizeof(DevState)) != cudaSuccess) { rc = fail(SDM_ERR_CUDA, "cudaMallocHost"); break; }
        memset(h->host_state, 0, sizeof(DevState));
        if (cudaMallocHost(&h->host_range, 16) != cudaSuccess) { rc = fail(SDM_ERR_CUDA, "cudaMallocHost"); break; }
        memset(h->host_range, 0, 16);
        if (h->shard_range.reserve(8) != cudaSuccess) { rc = fail(SDM_ERR_CUDA, "cudaMalloc range"); break; }
        if (h->state.reserve(1) != cudaSuccess) { rc = fail(SDM_ERR_CUDA, "cudaMalloc state"); break; }
        dev_fill(h->stream, h->state.p, 0, sizeof(DevState));
        cudaMemcpyToSymbolAsync(c_mc_packed, SDM_MC_PACKED_INIT, sizeof(SDM_MC_PACKED_INIT), 0, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyToSymbolAsync(c_mc_edgemask, SDM_MC_EDGEMASK_INIT, sizeof(SDM_MC_EDGEMASK_INIT), 0, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyToSymbolAsync(c_mc_ntri, SDM_MC_NTRI_INIT, sizeof(SDM_MC_NTRI_INIT), 0, cudaMemcpyHostToDevice, h->stream);
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = fail(SDM_ERR_CUDA, std::string("init: ") + cudaGetErrorString(cudaGetLastError())); break; }
        if (const char* e = getenv("SDM_SLACK")) { const float v = (float) atof(e); if (v > 0.0f && v <= 8.0f) h->slack_factor = v; }
        if (const char* e = getenv("SDM_PROJ_CHUNK")) { const int v = atoi(e); if (v >= 32 && v <= 65536) h->proj_chunk = (uint32_t) v & ~31u; }
        h->use_lists = getenv("SDM_NO_LISTS") == nullptr;
        h->use_lattice = getenv("SDM_NO_LATTICE") == nullptr;
        h->use_quick_orient = getenv("SDM_NO_QUICK_ORIENT") == nullptr;   // off: twelve evaluations for every triangle's orientation
        h->use_masks = getenv("SDM_NO_MASKS") == nullptr;   // off: every evaluation folds the whole table (the reference's own semantics)
        SdmPrimitive def[2];
        sdm_scene_default(def, 2);
        rc = sdm_set_scene(h, def, 2);
        if (rc) break;
        rc = ensure_capacity(h, 1u << 21);
    } while (0);
    if (rc) { sdm_destroy(h); return rc; }
    *out_handle = h;
    return SDM_OK;
}

void sdm_destroy(SdmHandle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    for (int b = 0; b < 2; b++) { if (h->ev_mesh_done[b]) cudaEventDestroy(h->ev_mesh_done[b]); if (h->ev_copy_done[b]) cudaEventDestroy(h->ev_copy_done[b]); }
    h->masks_fine.release(); h->masks_coarse.release(); h->cell_maybe.release();
    h->scene.release(); h->vox[0].release(); h->vox[1].release(); h->cases.release(); h->tri_off.release(); h->slot_ref.release();
    h->tri_uid.release(); h->first_slot.release(); h->wref.release(); h->first_bits.release(); h->first_prefix.release();
    h->tri_valid_bits.release(); h->orient_pending.release(); h->tri_prefix.release(); h->ustart.release(); h->upos.release(); h->unrm.release();
    for (int b = 0; b < 2; b++) { h->out_pos[b].release(); h->out_nrm[b].release(); h->out_idx[b].release(); }
    h->entry_uid.release(); h->vidx.release(); h->m27.release(); h->table1.release(); h->table2.release(); h->tiles.release();
    for (int i = 0; i < 2; i++) { h->vl[i].release(); h->vparent[i].release(); }
    h->urec.release(); h->tri_rec.release(); h->uesc.release(); h->stragglers.release(); h->state.release();
    if (h->host_state) cudaFreeHost(h->host_state);
    if (h->host_range) cudaFreeHost(h->host_range);
    if (h->host_scratch) cudaFreeHost(h->host_scratch);
    if (h->host_peer_local) cudaFreeHost(h->host_peer_local);
    for (void* m : h->ipc_mapped) if (m) cudaIpcCloseMemHandle(m);
    h->peer_block.release(); h->peer_local.release(); h->peer_frac.release(); h->peer_t0.release();
    h->shard_scratch.release();
    h->shard_range.release();
    h->render_target.release();
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// ---- probes ------------------------------------------------------------------------------------------
static int eval_common(SdmHandle* h, const float* points, uint32_t n, float* out, int width, int which, uint32_t* iters) {
    if (!h || !points || !out) return fail(SDM_ERR_INVALID, "null argument");
    if (n == 0) return SDM_OK;
    CK(cudaSetDevice(h->device));
    TempBuf<float> t_in, t_out;
    TempBuf<uint32_t> t_it;
    CK(t_in.alloc((size_t) n * 3));
    CK(t_out.alloc((size_t) n * width));
    if (iters) CK(t_it.alloc(n));
    float *d_in = t_in.p, *d_out = t_out.p;
    uint32_t* d_it = t_it.p;
    CK(cudaMemcpyAsync(d_in, points, (size_t) n * 12, cudaMemcpyHostToDevice, h->stream));
    const int grid = std::min<uint32_t>((n + 127) / 128, (uint32_t) h->num_sms * 8);
    {   // probes outside a remesh: masks over the last / default domain
        int rc = ensure_masks_any(h);
        if (rc) return rc;
    }
    if (which == 0) k_eval_sdf<<<grid, 128, smem_for(h, 128), h->stream>>>(h->scene.p, d_in, n, d_out, h->grid);
    else if (which == 1) k_eval_normal<<<grid, 128, smem_for(h, 128), h->stream>>>(h->scene.p, d_in, n, d_out, h->grid);
    else k_eval_project<<<grid, 128, smem_for(h, 128), h->stream>>>(h->scene.p, d_in, n, d_out, d_it, h->grid);
    h->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, (size_t) n * 4 * width, cudaMemcpyDeviceToHost, h->stream));
    if (iters) CK(cudaMemcpyAsync(iters, d_it, (size_t) n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SDM_OK;
}
int sdm_eval_sdf(SdmHandle* h, const float* points, uint32_t n, float* out_sd) { return eval_common(h, points, n, out_sd, 1, 0, nullptr); }
int sdm_eval_normal(SdmHandle* h, const float* points, uint32_t n, float* out_normals) { return eval_common(h, points, n, out_normals, 3, 1, nullptr); }
int sdm_eval_project(SdmHandle* h, const float* points, uint32_t n, float* out_points, uint32_t* out_iters) {
    return eval_common(h, points, n, out_points, 3, 2, out_iters);
}

// ---- host-buffer surface (CudaHandler semantics) --------------------------------------------------------
int sdm_create_voxel_field(const SdmParams* params, SdmVoxelField* out_field) {
    if (!out_field) return fail(SDM_ERR_INVALID, "null out_field");
    SdmParams p { SDM_MESH_GENERATION_BB_SIZE, SDM_MESH_GENERATION_INIT_FACTOR, 0 };
    if (params) p = *params;
    int rc = check_params(p);
    if (rc) return rc;
    const uint32_t N = p.init_factor;
    const float size = p.bb_size / (float) N;   // src/cuda/mod.rs:106
    const size_t n = (size_t) N * N * N;
    SdmPoint* v = (SdmPoint*) malloc(std::max<size_t>(n, 1) * sizeof(SdmPoint));
    if (!v) return fail(SDM_ERR_INVALID, "out of host memory");
    size_t o = 0;
    for (uint32_t x = 0; x < N; x++)          // x outer, z inner (:110-119)
        for (uint32_t y = 0; y < N; y++)
            for (uint32_t z = 0; z < N; z++, o++) {
                v[o].x = (float) x * size - p.bb_size / 2.0f;
                v[o].y = (float) y * size - p.bb_size / 2.0f;
                v[o].z = (float) z * size - p.bb_size / 2.0f;
            }
    out_field->voxel_size = SdmPoint { size, size, size };
    out_field->voxels = v;
    out_field->voxel_count = (unsigned int) n;
    return SDM_OK;
}
void sdm_voxel_field_free(SdmVoxelField* field) {
    if (field && field->voxels) { free(field->voxels); field->voxels = nullptr; field->voxel_count = 0; }
}

int sdm_field_upload(SdmHandle* h, const SdmVoxelField* field) {
    if (!h || !field || (!field->voxels && field->voxel_count)) return fail(SDM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    uint32_t want = 0;
    int rc = capacity_for(h, field->voxel_count, &want);
    if (rc) return rc;
    rc = ensure_capacity(h, want);
    if (rc) return rc;
    rc = reset_state(h);
    if (rc) return rc;
    rc = ensure_masks_any(h);   // an uploaded list carries no domain: masks over the last / default cube
    if (rc) return rc;
    if (field->voxel_count)
        CK(cudaMemcpyAsync(h->vox[0].p, field->voxels, (size_t) field->voxel_count * 12, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(&h->state.p->level_count[0], &field->voxel_count, 4, cudaMemcpyHostToDevice, h->stream));
    h->cur = 0; h->level = 0;
    h->cases_for_level = -1;
    h->voxel_size[0] = field->voxel_size.x; h->voxel_size[1] = field->voxel_size.y; h->voxel_size[2] = field->voxel_size.z;
    h->have_field = true; h->mesh_valid = false;
    h->lists_level = -1;          // an uploaded list has no ancestors: cell masks
    h->lattice.enabled = 0;       // ... and no known origin: generic vertex keys
    CK(cudaStreamSynchronize(h->stream));   // the host list may be freed by the caller right after
    return SDM_OK;
}

int sdm_field_reset(SdmHandle* h, const SdmParams* params) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    SdmParams p { SDM_MESH_GENERATION_BB_SIZE, SDM_MESH_GENERATION_INIT_FACTOR, 0 };
    if (params) p = *params;
    int rc = check_params(p);
    if (rc) return rc;
    CK(cudaSetDevice(h->device));
    uint32_t want = 0;
    rc = capacity_for(h, (uint64_t) p.init_factor * p.init_factor * p.init_factor, &want);
    if (rc) return rc;
    rc = ensure_capacity(h, want);
    if (rc) return rc;
    rc = reset_state(h);
    if (rc) return rc;
    return enqueue_init_field(h, p);
}

int sdm_field_refine(SdmHandle* h, uint32_t* out_count) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    if (!h->have_field) return fail(SDM_ERR_STATE, "no device field: call sdm_field_reset / sdm_field_upload first");
    CK(cudaSetDevice(h->device));
    // a level can at most multiply the list by 8; grow-and-retry keeps the reference's "always fits" behaviour
    for (int attempt = 0; attempt < 8; attempt++) {
        int rc = enqueue_refine(h);
        if (rc) return rc;
        uint32_t flags = 0;
        rc = fetch_state(h, &flags);
        if (rc) return rc;
        if (!(flags & ERR_VOXEL_CAP)) {
            if (out_count) *out_count = h->host_state->level_count[h->level];
            fill_stats(h, false);
            return SDM_OK;
        }
        // overflow: download the parent list, grow, re-upload, retry
        const int parent_level = h->level - 1;
        const uint32_t n_parent = h->host_state->level_count[parent_level];
        std::vector<float> parents((size_t) n_parent * 3);
        CK(cudaMemcpy(parents.data(), h->vox[h->cur ^ 1].p, parents.size() * 4, cudaMemcpyDeviceToHost));
        const float ps[3] = { h->voxel_size[0] * 2.0f, h->voxel_size[1] * 2.0f, h->voxel_size[2] * 2.0f };
        SdmVoxelField f { SdmPoint { ps[0], ps[1], ps[2] }, (SdmPoint*) parents.data(), n_parent };
        if (grown(h->cap_vox) == h->cap_vox) break;
        rc = ensure_capacity(h, grown(h->cap_vox));
        if (rc) return rc;
        rc = sdm_field_upload(h, &f);
        if (rc) return rc;
    }
    return fail(SDM_ERR_CAPACITY, "voxel list does not fit in device memory");
}

int sdm_field_count(SdmHandle* h, uint32_t* out_count, SdmPoint* out_voxel_size) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    if (!h->have_field) return fail(SDM_ERR_STATE, "no device field");
    CK(cudaSetDevice(h->device));
    uint32_t flags = 0;
    int rc = fetch_state(h, &flags);
    if (rc) return rc;
    if (out_count) *out_count = h->host_state->level_count[h->level];
    if (out_voxel_size) *out_voxel_size = SdmPoint { h->voxel_size[0], h->voxel_size[1], h->voxel_size[2] };
    return SDM_OK;
}

int sdm_field_download(SdmHandle* h, SdmPoint* out_voxels, uint32_t capacity) {
    uint32_t n = 0;
    int rc = sdm_field_count(h, &n, nullptr);
    if (rc) return rc;
    if (n > capacity) return fail(SDM_ERR_INVALID, "capacity too small");
    if (n) CK(cudaMemcpy(out_voxels, h->vox[h->cur].p, (size_t) n * 12, cudaMemcpyDeviceToHost));
    return SDM_OK;
}

static int run_mesh(SdmHandle* h, SdmMesh* out_mesh, bool timed) {
    for (int attempt = 0; attempt < 8; attempt++) {
        if (timed) CK(cudaEventRecord(h->ev0, h->stream));
        int rc = enqueue_mesh(h);
        if (rc) return rc;
        if (timed) CK(cudaEventRecord(h->ev1, h->stream));
        uint32_t flags = 0;
        rc = fetch_state(h, &flags);
        if (rc) return rc;
        if (!flags) {
            if (timed) { float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->stats.last_gpu_ms = ms; }
            h->mesh_valid = true;
            fill_stats(h, true);
            adapt_table1(h);
            if (out_mesh) mesh_view(h, out_mesh);
            return SDM_OK;
        }
        if (only_table1_overflow(h, flags)) {
            CK(dev_fill(h->stream, &h->state.p->error_flags, 0, 4));
            continue;
        }
        // a mesh-stage capacity was exceeded: keep the field (download / grow / upload) and retry
        const uint32_t n = h->host_state->level_count[h->level];
        std::vector<float> list((size_t) n * 3);
        if (n) CK(cudaMemcpy(list.data(), h->vox[h->cur].p, list.size() * 4, cudaMemcpyDeviceToHost));
        SdmVoxelField f { SdmPoint { h->voxel_size[0], h->voxel_size[1], h->voxel_size[2] }, (SdmPoint*) list.data(), n };
        if (grown(h->cap_vox) == h->cap_vox) break;
        rc = ensure_capacity(h, grown(h->cap_vox));
        if (rc) return rc;
        rc = sdm_field_upload(h, &f);
        if (rc) return rc;
    }
    return fail(SDM_ERR_CAPACITY, "mesh does not fit in device memory");
}

int sdm_field_to_mesh(SdmHandle* h, SdmMesh* out_mesh) {
    if (!h || !out_mesh) return fail(SDM_ERR_INVALID, "null argument");
    if (!h->have_field) return fail(SDM_ERR_STATE, "no device field");
    CK(cudaSetDevice(h->device));
    return run_mesh(h, out_mesh, true);
}

int sdm_field_cases(SdmHandle* h, uint8_t* out_cases, uint32_t capacity) {
    if (!h || !out_cases) return fail(SDM_ERR_INVALID, "null argument");
    if (!h->mesh_valid) return fail(SDM_ERR_STATE, "no mesh: call sdm_field_to_mesh / sdm_remesh first");
    const uint32_t n = h->host_state->level_count[h->level];
    if (n > capacity) return fail(SDM_ERR_INVALID, "capacity too small");
    CK(cudaSetDevice(h->device));
    if (n) CK(cudaMemcpy(out_cases, h->cases.p, n, cudaMemcpyDeviceToHost));
    return SDM_OK;
}

int sdm_field_triangle_soup(SdmHandle* h, SdmTriangle* out_triangles, uint32_t capacity) {
    if (!h || !out_triangles) return fail(SDM_ERR_INVALID, "null argument");
    if (!h->mesh_valid) return fail(SDM_ERR_STATE, "no mesh: call sdm_field_to_mesh / sdm_remesh first");
    const uint32_t n = h->host_state->level_count[h->level];
    if ((uint64_t) n * 5 > capacity) return fail(SDM_ERR_INVALID, "capacity too small");
    if (n == 0) return SDM_OK;
    CK(cudaSetDevice(h->device));
    TempBuf<float> tmp;
    CK(tmp.alloc((size_t) n * 5 * 18));
    float* d = tmp.p;
    k_soup<<<h->g_light, 256, 0, h->stream>>>(h->state.p, h->level, h->cases.p, h->tri_off.p, h->tri_uid.p, h->upos.p, h->unrm.p, d);
    h->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_triangles, d, (size_t) n * 5 * sizeof(SdmTriangle), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SDM_OK;
}

int sdm_remesh(SdmHandle* h, const SdmParams* params, SdmMesh* out_mesh) {
    if (!h || !out_mesh) return fail(SDM_ERR_INVALID, "null argument");
    SdmParams p { SDM_MESH_GENERATION_BB_SIZE, SDM_MESH_GENERATION_INIT_FACTOR, 0 };
    if (params) p = *params;
    int rc = check_params(p);
    if (rc) return rc;
    CK(cudaSetDevice(h->device));
    uint32_t want = 0;
    rc = capacity_for(h, (uint64_t) p.init_factor * p.init_factor * p.init_factor, &want);
    if (rc) return rc;
    static const bool trace = getenv("SDM_TRACE") != nullptr;   // developer switch: host-side phase times of sdm_remesh on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    for (int attempt = 0; attempt < 10; attempt++) {
        const auto t_begin = now();
        rc = ensure_capacity(h, want);
        if (rc) return rc;
        prof_begin(h);
        CK(cudaEventRecord(h->ev0, h->stream));
        rc = reset_state(h);
        if (rc) return rc;
        mark(h, "start");
        NvtxRange nv(h, "sdm_remesh");
        rc = enqueue_init_field(h, p);
        if (rc) return rc;
        // the final voxel size is known: one inflation (slack of the mesh stage) for the list regions of all levels
        h->delta_override = h->slack_factor * ldexpf(p.bb_size / (float) p.init_factor, -(int) p.levels);
        for (uint32_t l = 0; l < p.levels && !rc; l++) rc = enqueue_refine(h, l + 1 == p.levels);
        h->delta_override = 0.0f;
        if (rc) return rc;
        rc = enqueue_mesh(h);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev1, h->stream));
        const double t_enqueue = ms_since(t_begin);
        uint32_t flags = 0;
        rc = fetch_state(h, &flags);
        if (rc) return rc;
        if (!flags) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev0, h->ev1);
            if (trace) fprintf(stderr, "sdm_remesh: enqueue %.3f ms, total %.3f ms, gpu %.3f ms\n", t_enqueue, ms_since(t_begin), ms);
            h->stats.last_gpu_ms = ms;
            h->mesh_valid = true;
            fill_stats(h, true);
            prof_end(h);
            adapt_table1(h);
            mesh_view(h, out_mesh);
            return SDM_OK;
        }
        if (only_table1_overflow(h, flags)) continue;
        want = grown(h->cap_vox);
        if (want == h->cap_vox) break;
    }
    return fail(SDM_ERR_CAPACITY, "remesh does not fit in device memory");
}

int sdm_mesh_download(SdmHandle* h, const SdmMesh* m, float* positions, float* normals, uint32_t* indices) {
    if (!h || !m) return fail(SDM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    if (m->vertex_count && positions) CK(cudaMemcpyAsync(positions, m->positions, (size_t) m->vertex_count * 12, cudaMemcpyDeviceToHost, h->stream));
    if (m->vertex_count && normals) CK(cudaMemcpyAsync(normals, m->normals, (size_t) m->vertex_count * 12, cudaMemcpyDeviceToHost, h->stream));
    if (m->triangle_count && indices) CK(cudaMemcpyAsync(indices, m->indices, (size_t) m->triangle_count * 12, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SDM_OK;
}

// Asynchronous hand-off to the host: copies on the handle's copy stream, ordered after the weld that produced `m`; the next
// sdm_remesh may be issued immediately (it writes the other output set).  Host memory should be pinned.
int sdm_mesh_download_async(SdmHandle* h, const SdmMesh* m, float* positions, float* normals, uint32_t* indices) {
    if (!h || !m) return fail(SDM_ERR_INVALID, "null argument");
    if (!m->on_device || m->reserved < 0 || m->reserved > 7) return fail(SDM_ERR_INVALID, "not a device mesh of this library");
    CK(cudaSetDevice(h->device));
    // output set of the indices; a mesh assembled by sdm_shard_fixup keeps its vertices in the other set (reserved = 4 | vset << 1 | iset)
    const int b = m->reserved & 1, vb = (m->reserved & 4) ? ((m->reserved >> 1) & 1) : b;
    cudaStream_t cs = h->copy_stream;
    CK(cudaStreamWaitEvent(cs, h->ev_mesh_done[b], 0));
    if (m->vertex_count && positions) CK(cudaMemcpyAsync(positions, m->positions, (size_t) m->vertex_count * 12, cudaMemcpyDeviceToHost, cs));
    if (m->vertex_count && normals) CK(cudaMemcpyAsync(normals, m->normals, (size_t) m->vertex_count * 12, cudaMemcpyDeviceToHost, cs));
    if (m->triangle_count && indices) CK(cudaMemcpyAsync(indices, m->indices, (size_t) m->triangle_count * 12, cudaMemcpyDeviceToHost, cs));
    CK(cudaEventRecord(h->ev_copy_done[b], cs));
    if (vb != b) CK(cudaEventRecord(h->ev_copy_done[vb], cs));   // the next weld into either set waits for this download
    (void) cudaStreamQuery(cs);   // push the copies to the device now
    return SDM_OK;
}
int sdm_mesh_download_wait(SdmHandle* h) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->copy_stream));
    return SDM_OK;
}

uint64_t sdm_hash_bytes(const void* data, size_t bytes) {
    const unsigned char* b = (const unsigned char*) data;
    uint64_t x = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < bytes; i++) { x ^= b[i]; x *= 0x100000001b3ull; }
    return x;
}

// ---- OBJ writer (src/renderer/mod.rs:204 / obj crate `ObjData::save`) ---------------------------------------------------------
namespace {
// Rust's `{}` for f32: shortest digits that read back as the same f32 (ties: the candidate closest to the value), never
// in exponent notation.  Digits: the correctly rounded p-digit decimal for the smallest p (1..9) that round-trips.
int format_f32_rust(float v, char* out /* >= 64 */) {
    if (std::isnan(v)) return snprintf(out, 64, "NaN");
    if (std::isinf(v)) return snprintf(out, 64, v > 0 ? "inf" : "-inf");
    if (v == 0.0f) return snprintf(out, 64, std::signbit(v) ? "-0" : "0");
    char buf[48];
    char digits[16] = { 0 };
    int nd = 0, exp10 = 0;
    const double av = std::fabs((double) v);
    for (int p = 1; p <= 9 && nd == 0; p++) {
        // the correctly rounded p-digit decimal and its two p-digit neighbours: at a power of two the interval of decimals that
        // read back as v is lopsided, and the shortest representation can be a neighbour of the rounded one
        snprintf(buf, sizeof buf, "%.*e", p - 1, av);
        long long D = 0;
        const char* q = buf;
        for (; *q && *q != 'e'; q++) if (*q >= '0' && *q <= '9') D = D * 10 + (*q - '0');
        const int e10 = atoi(q + 1);
        double best_err = 0.0;
        long long best = -1;
        int best_e = 0;
        for (int k = 0; k < 3; k++) {                                  // the rounded decimal first: it wins ties
            const long long c = k == 0 ? D : (k == 1 ? D - 1 : D + 1);
            if (c <= 0) continue;
            long long cd = c;
            int ce = e10;
            long long lim = 1;
            for (int i = 0; i < p; i++) lim *= 10;
            if (cd >= lim) { cd = lim / 10; ce += 1; }                 // 999 + 1 -> 100 with the next exponent
            else if (cd < lim / 10) { cd = lim - 1; ce -= 1; }         // 100 - 1 -> 999 with the previous exponent
            snprintf(buf, sizeof buf, "%llde%d", cd, ce - (p - 1));
            if (strtof(buf, nullptr) != std::fabs(v)) continue;
            const double err = std::fabs(strtod(buf, nullptr) - av);
            if (best < 0 || err < best_err * (1.0 - 1e-7)) { best = cd; best_err = err; best_e = ce; }
        }
        if (best >= 0) {
            nd = snprintf(digits, sizeof digits, "%lld", best);
            exp10 = best_e;
        }
    }
    while (nd > 1 && digits[nd - 1] == '0') nd--;   // "1.0e0" style candidates never occur for p = shortest, but stay safe
    char* o = out;
    if (v < 0) *o++ = '-';
    if (exp10 < 0) {                                // 0.000ddd
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -exp10 - 1; i++) *o++ = '0';
        for (int i = 0; i < nd; i++) *o++ = digits[i];
    } else if (exp10 >= nd - 1) {                   // ddd000
        for (int i = 0; i < nd; i++) *o++ = digits[i];
        for (int i = 0; i < exp10 - (nd - 1); i++) *o++ = '0';
    } else {                                        // dd.ddd
        for (int i = 0; i <= exp10; i++) *o++ = digits[i];
        *o++ = '.';
        for (int i = exp10 + 1; i < nd; i++) *o++ = digits[i];
    }
    *o = 0;
    return (int) (o - out);
}
}  // namespace

int sdm_mesh_save_obj(SdmHandle* h, const SdmMesh* m, const char* path) {
    if (!m || !path) return fail(SDM_ERR_INVALID, "null argument");
    std::vector<float> pos, nrm;
    std::vector<uint32_t> idx;
    const float *P = m->positions, *N = m->normals;
    const uint32_t* I = m->indices;
    if (m->on_device) {
        if (!h) return fail(SDM_ERR_INVALID, "a device mesh needs its handle");
        pos.resize((size_t) m->vertex_count * 3); nrm.resize((size_t) m->vertex_count * 3); idx.resize((size_t) m->triangle_count * 3);
        int rc = sdm_mesh_download(h, m, pos.data(), nrm.data(), idx.data());
        if (rc) return rc;
        P = pos.data(); N = nrm.data(); I = idx.data();
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SDM_ERR_INVALID, std::string("cannot open ") + path);
    std::vector<char> wb(1 << 20);
    setvbuf(f, wb.data(), _IOFBF, wb.size());
    char a[64], b[64], c[64];
    for (uint32_t i = 0; i < m->vertex_count; i++) {
        format_f32_rust(P[3 * (size_t) i], a); format_f32_rust(P[3 * (size_t) i + 1], b); format_f32_rust(P[3 * (size_t) i + 2], c);
        fprintf(f, "v %s %s %s\n", a, b, c);
    }
    fputs("vt 0 0\n", f);                                    // texture: vec![[0.0, 0.0]] (src/cuda/mod.rs:306)
    for (uint32_t i = 0; i < m->vertex_count; i++) {
        format_f32_rust(N[3 * (size_t) i], a); format_f32_rust(N[3 * (size_t) i + 1], b); format_f32_rust(N[3 * (size_t) i + 2], c);
        fprintf(f, "vn %s %s %s\n", a, b, c);
    }
    fputs("o default\ng default\n", f);                      // one object, one group, both named "default" (:308-311)
    for (uint32_t t = 0; t < m->triangle_count; t++) {       // IndexTuple(idx, Some(0), Some(idx)) (:320), 1-based in the file
        const uint32_t x = I[3 * (size_t) t] + 1, y = I[3 * (size_t) t + 1] + 1, z = I[3 * (size_t) t + 2] + 1;
        fprintf(f, "f %u/1/%u %u/1/%u %u/1/%u\n", x, x, y, y, z, z);
    }
    const bool bad = ferror(f) != 0;
    if (fclose(f) != 0 || bad) return fail(SDM_ERR_INVALID, std::string("write error on ") + path);
    return SDM_OK;
}

int sdm_refine_voxel_field(SdmHandle* h, SdmVoxelField* field) {
    if (!h || !field) return fail(SDM_ERR_INVALID, "null argument");
    if (field->voxel_count == 0) return SDM_OK;   // src/cuda/mod.rs:137: size is NOT halved for an empty field
    int rc = sdm_field_upload(h, field);
    if (rc) return rc;
    uint32_t n = 0;
    rc = sdm_field_refine(h, &n);
    if (rc) return rc;
    SdmPoint* v = (SdmPoint*) malloc(std::max<size_t>(n, 1) * sizeof(SdmPoint));
    if (!v) return fail(SDM_ERR_INVALID, "out of host memory");
    rc = sdm_field_download(h, v, n);
    if (rc) { free(v); return rc; }
    free(field->voxels);
    field->voxels = v;
    field->voxel_count = n;
    field->voxel_size = SdmPoint { h->voxel_size[0], h->voxel_size[1], h->voxel_size[2] };
    return SDM_OK;
}

int sdm_voxel_field_to_mesh(SdmHandle* h, const SdmVoxelField* field, SdmMesh* out_mesh) {
    if (!h || !field || !out_mesh) return fail(SDM_ERR_INVALID, "null argument");
    memset(out_mesh, 0, sizeof(*out_mesh));
    if (field->voxel_count == 0) return SDM_OK;   // empty mesh (src/cuda/mod.rs:327-345)
    int rc = sdm_field_upload(h, field);
    if (rc) return rc;
    SdmMesh dm;
    rc = run_mesh(h, &dm, true);
    if (rc) return rc;
    out_mesh->vertex_count = dm.vertex_count;
    out_mesh->triangle_count = dm.triangle_count;
    out_mesh->positions = (float*) malloc(std::max<size_t>(dm.vertex_count, 1) * 12);
    out_mesh->normals = (float*) malloc(std::max<size_t>(dm.vertex_count, 1) * 12);
    out_mesh->indices = (uint32_t*) malloc(std::max<size_t>(dm.triangle_count, 1) * 12);
    if (!out_mesh->positions || !out_mesh->normals || !out_mesh->indices) { sdm_mesh_free(out_mesh); return fail(SDM_ERR_INVALID, "out of host memory"); }
    return sdm_mesh_download(h, &dm, out_mesh->positions, out_mesh->normals, out_mesh->indices);
}

void sdm_mesh_free(SdmMesh* mesh) {
    if (!mesh || mesh->on_device) return;
    free(mesh->positions); free(mesh->normals); free(mesh->indices);
    memset(mesh, 0, sizeof(*mesh));
}

// ---- shards ----------------------------------------------------------------------------------------------
int sdm_shard_remesh(SdmHandle* h, const SdmParams* params, uint32_t split_level, uint32_t shard_index, uint32_t shard_count,
                     SdmShardInfo* out_info) {
    if (!h || !out_info) return fail(SDM_ERR_INVALID, "null argument");
    SdmParams p { SDM_MESH_GENERATION_BB_SIZE, SDM_MESH_GENERATION_INIT_FACTOR, 0 };
    if (params) p = *params;
    int rc = check_params(p);
    if (rc) return rc;
    if (shard_count == 0 || shard_index >= shard_count) return fail(SDM_ERR_INVALID, "bad shard index / count");
    if (split_level > p.levels) split_level = p.levels;
    CK(cudaSetDevice(h->device));
    uint32_t want = 0;
    rc = capacity_for(h, (uint64_t) p.init_factor * p.init_factor * p.init_factor, &want);
    if (rc) return rc;
    for (int attempt = 0; attempt < 10; attempt++) {
        rc = ensure_capacity(h, want);
        if (rc) return rc;
        prof_begin(h);
        CK(cudaEventRecord(h->ev0, h->stream));
        rc = reset_state(h);
        if (rc) return rc;
        mark(h, "start");
        rc = enqueue_init_field(h, p);
        if (rc) return rc;
        h->delta_override = h->slack_factor * ldexpf(p.bb_size / (float) p.init_factor, -(int) p.levels);
        for (uint32_t l = 0; l < split_level && !rc; l++) rc = enqueue_refine(h);   // redundantly on every rank: the coarse levels are tiny
        if (rc) { h->delta_override = 0.0f; return rc; }
        rc = enqueue_take_shard(h, shard_index, shard_count);
        if (rc) { h->delta_override = 0.0f; return rc; }
        for (uint32_t l = split_level; l < p.levels && !rc; l++) rc = enqueue_refine(h, l + 1 == p.levels);
        h->delta_override = 0.0f;
        if (rc) return rc;
        rc = enqueue_mesh_local(h, true);   // weld keys too: sdm_shard_local_weld follows (the root-weld fallback re-inserts)
        if (rc) return rc;
        CK(cudaEventRecord(h->ev1, h->stream));
        CK(cudaMemcpyAsync(h->host_range, h->shard_range.p, 12, cudaMemcpyDeviceToHost, h->stream));
        uint32_t flags = 0;
        rc = fetch_state(h, &flags);
        if (rc) return rc;
        if (!flags) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev0, h->ev1);
            h->stats.last_gpu_ms = ms;
            h->mesh_valid = false;   // no welded mesh yet
            fill_stats(h, true);
            prof_end(h);
            h->own_tris = h->host_state->n_tris_raw;
            h->own_uniq = h->host_state->n_uniq;
            adapt_table1(h);
            h->shard_index = shard_index;
            out_info->shard_index = shard_index; out_info->shard_count = shard_count; out_info->split_level = split_level;
            out_info->voxel_begin = h->host_range[0]; out_info->voxel_end = h->host_range[1]; out_info->split_total = h->host_range[2];
            out_info->final_voxels = h->host_state->level_count[h->level];
            out_info->unique_vertices = h->own_uniq; out_info->raw_triangles = h->own_tris;
            return SDM_OK;
        }
        if (only_table1_overflow(h, flags)) continue;
        want = grown(h->cap_vox);
        if (want == h->cap_vox) break;
    }
    return fail(SDM_ERR_CAPACITY, "shard does not fit in device memory");
}

int sdm_shard_buffers(SdmHandle* h, SdmShardBuffers* out) {
    if (!h || !out) return fail(SDM_ERR_INVALID, "null argument");
    out->positions = h->upos.p; out->normals = h->unrm.p; out->triangle_vertex_ids = h->tri_uid.p;
    out->capacity_vertices = h->cap_uniq; out->capacity_triangles = h->cap_tris;
    return SDM_OK;
}

int sdm_shard_prepare_send(SdmHandle* h, uint32_t vertex_offset) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    k_shard_prepare_send<<<h->g_light, 256, 0, h->stream>>>(h->state.p, h->tri_uid.p, h->tri_valid_bits.p, vertex_offset);
    h->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));   // the buffers are handed to NCCL on another stream next
    return SDM_OK;
}

int sdm_shard_reserve(SdmHandle* h, uint32_t total_vertices, uint32_t total_triangles) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    uint32_t want = h->cap_vox;
    while ((uint64_t) want * 2 < total_vertices || (uint64_t) want * 3 < total_triangles) {
        const uint32_t g = grown(want);
        if (g == want) return fail(SDM_ERR_CAPACITY, "merged mesh too large");
        want = g;
    }
    if (want == h->cap_vox) return SDM_OK;
    CK(cudaStreamSynchronize(h->stream));
    // keep the local shard's vertices / triangles: detach, re-allocate everything, copy back
    float* old_pos = h->upos.p; float* old_nrm = h->unrm.p; uint32_t* old_uid = h->tri_uid.p; uint32_t* old_bits = h->tri_valid_bits.p;
    h->upos.p = nullptr; h->upos.n = 0; h->unrm.p = nullptr; h->unrm.n = 0; h->tri_uid.p = nullptr; h->tri_uid.n = 0;
    h->tri_valid_bits.p = nullptr; h->tri_valid_bits.n = 0;
    int rc = ensure_capacity(h, want);
    if (rc == SDM_OK) {
        cudaMemcpyAsync(h->upos.p, old_pos, (size_t) h->own_uniq * 12, cudaMemcpyDeviceToDevice, h->stream);
        cudaMemcpyAsync(h->unrm.p, old_nrm, (size_t) h->own_uniq * 12, cudaMemcpyDeviceToDevice, h->stream);
        cudaMemcpyAsync(h->tri_uid.p, old_uid, (size_t) h->own_tris * 12, cudaMemcpyDeviceToDevice, h->stream);
        cudaMemcpyAsync(h->tri_valid_bits.p, old_bits, ((size_t) h->own_tris + 31) / 32 * 4, cudaMemcpyDeviceToDevice, h->stream);
        cudaStreamSynchronize(h->stream);
    }
    cudaFree(old_pos); cudaFree(old_nrm); cudaFree(old_uid); cudaFree(old_bits);
    return rc;
}

static int ensure_shard_scratch(SdmHandle* h) {
    CK(h->shard_scratch.reserve(128));
    if (!h->host_scratch) CK(cudaMallocHost(&h->host_scratch, 128 * sizeof(uint32_t)));
    return SDM_OK;
}
static float ord_to_float(uint32_t o) { const int32_t i = (int32_t) o; const int32_t b = i ^ ((i >> 31) & 0x7fffffff); float f; memcpy(&f, &b, 4); return f; }

int sdm_shard_local_weld(SdmHandle* h, SdmShardWeld* out) {
    if (!h || !out) return fail(SDM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    int rc = ensure_shard_scratch(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    CK(cudaEventRecord(h->ev0, s));
    const int b = h->out_sel;
    rc = enqueue_weld(h, true);
    if (rc) return rc;
    k_shard_scratch_init<<<1, 32, 0, s>>>(h->shard_scratch.p);
    k_shard_xrange<<<h->g_light, 256, 0, s>>>(h->state.p, h->out_pos[b].p, h->shard_scratch.p);
    h->stats.kernel_launches += 2;
    CK(cudaEventRecord(h->ev1, s));
    CK(cudaMemcpyAsync(h->host_scratch, h->shard_scratch.p, 16, cudaMemcpyDeviceToHost, s));
    uint32_t flags = 0;
    rc = fetch_state(h, &flags);
    if (rc) return rc;
    if (flags) return fail(SDM_ERR_CAPACITY, "weld table overflow");
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->stats.last_gpu_ms = ms;
    h->welded_set = b; h->welded_vset = b;
    h->out_sel ^= 1;   // like mesh_view: the next weld writes the other set
    h->welded_v = h->host_state->n_verts_out; h->welded_t = h->host_state->n_tris_out;
    out->vertices = h->welded_v;
    out->triangles = h->welded_t;
    out->nonfinite = h->host_scratch[2] + (h->welded_v >= (1u << 24) ? 1u : 0u);   // 24-bit local indices in the key rows
    out->min_x = ord_to_float(h->host_scratch[0]);
    out->max_x = ord_to_float(h->host_scratch[1]);
    if (h->host_scratch[0] == 0x7fffffffu) { out->min_x = 1.0f; out->max_x = 0.0f; }
    return SDM_OK;
}

int sdm_shard_boundary_keys(SdmHandle* h, const float* lo, const float* hi, uint32_t interval_count, uint32_t** out_rows_device, uint32_t* out_count) {
    if (!h || !out_rows_device || !out_count || (interval_count && (!lo || !hi))) return fail(SDM_ERR_INVALID, "null argument");
    if (interval_count > 32) return fail(SDM_ERR_INVALID, "at most 32 intervals");
    if (h->shard_index > 255) return fail(SDM_ERR_INVALID, "at most 256 shards");
    CK(cudaSetDevice(h->device));
    ShardIntervals iv {};
    iv.count = interval_count;
    for (uint32_t i = 0; i < interval_count; i++) { iv.lo[i] = lo[i]; iv.hi[i] = hi[i]; }
    cudaStream_t s = h->stream;
    const uint32_t cap_rows = h->table_entries;   // the vertex table is free after k_orient: its 16-byte entries hold the rows
    CK(dev_fill(s, h->shard_scratch.p + 3, 0, 4));   // the candidate counter
    k_shard_boundary_keys<<<h->g_light, 256, 0, s>>>(h->state.p, h->out_pos[h->welded_vset].p, iv, h->shard_index, h->table1.p, cap_rows, h->shard_scratch.p);
    h->stats.kernel_launches++;
    CK(cudaMemcpyAsync(h->host_scratch, h->shard_scratch.p, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h->host_scratch[3] > cap_rows) return fail(SDM_ERR_CAPACITY, "too many boundary candidates");
    *out_rows_device = reinterpret_cast<uint32_t*>(h->table1.p);
    *out_count = h->host_scratch[3];
    return SDM_OK;
}

int sdm_shard_key_scratch(SdmHandle* h, uint32_t rows, uint32_t** out_rows_device) {
    if (!h || !out_rows_device) return fail(SDM_ERR_INVALID, "null argument");
    if (rows > h->table_entries) return fail(SDM_ERR_CAPACITY, "too many boundary candidates");
    *out_rows_device = reinterpret_cast<uint32_t*>(h->table1.p);
    return SDM_OK;
}

int sdm_shard_resolve(SdmHandle* h, const uint32_t* rows_device, uint32_t total, const uint32_t* vertex_counts, uint32_t shards, uint32_t* out_removed) {
    if (!h || !vertex_counts || !out_removed || (total && !rows_device)) return fail(SDM_ERR_INVALID, "null argument");
    if (shards == 0 || shards > 32) return fail(SDM_ERR_INVALID, "1..32 shards");
    CK(cudaSetDevice(h->device));
    int rc = ensure_shard_scratch(h);
    if (rc) return rc;
    ShardOffsets& so = h->res_offsets;
    so = ShardOffsets {};
    so.count = shards;
    uint64_t vsum = 0;
    for (uint32_t s = 0; s < shards; s++) { so.voff[s] = (uint32_t) vsum; vsum += vertex_counts[s]; }
    so.voff[shards] = (uint32_t) vsum;
    const uint32_t nwords = (uint32_t) (vsum / 32 + 2);
    const uint32_t entries = std::min<uint32_t>(pow2_at_least(std::max<uint64_t>((uint64_t) total * 2, 1024)), h->table_entries);
    if (vsum > h->cap_uniq || nwords > h->first_bits.n || nwords > h->first_prefix.n || total > h->wref.n || (uint64_t) total * 2 > h->slot_ref.n ||
        (uint64_t) entries * 4 < (uint64_t) total * 5)
        return fail(SDM_ERR_CAPACITY, "call sdm_shard_reserve_welded with the totals first");
    cudaStream_t st = h->stream;
    uint32_t* sc = h->shard_scratch.p;   // [32..64) duplicates per shard, [64..96) cursors, [96..128) global offsets, [4] errors
    CK(dev_fill(st, sc, 0, 128 * 4));
    CK(dev_fill(st, h->first_bits.p, 0, (size_t) nwords * 4));
    if (total) {
        CK(dev_fill(st, h->table2.p, 0xFF, (size_t) entries * 16));
        const uint4* rows = reinterpret_cast<const uint4*>(rows_device);
        k_res_insert<<<h->g_light, 256, 0, st>>>(rows, total, h->table2.p, entries - 1, h->wref.p, sc + 4);
        k_res_mark<<<h->g_light, 256, 0, st>>>(rows, total, h->table2.p, h->wref.p, so, h->first_bits.p, sc + 32);
    }
    k_scan_bits_1block<<<1, 1024, 0, st>>>(h->first_bits.p, h->first_prefix.p, nwords);
    CK(cudaMemcpyAsync(h->host_scratch, sc, 128 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h->host_scratch[4]) return fail(SDM_ERR_CAPACITY, "key table overflow in sdm_shard_resolve");
    uint32_t psum = 0;
    for (uint32_t s = 0; s < shards; s++) { out_removed[s] = h->host_scratch[32 + s]; so.poff[s] = psum; psum += out_removed[s]; }
    so.poff[shards] = psum;
    k_res_pairs<<<h->g_light, 256, 0, st>>>(reinterpret_cast<const uint4*>(rows_device), total, h->table2.p, h->wref.p, so, h->first_bits.p, h->first_prefix.p,
                                            sc + 64, reinterpret_cast<uint2*>(h->slot_ref.p), sc + 96);
    h->stats.kernel_launches += 7;
    CK(cudaGetLastError());
    return SDM_OK;
}

int sdm_shard_fixup(SdmHandle* h, const uint32_t* triangle_counts, SdmMesh* out_mesh) {
    if (!h || !triangle_counts || !out_mesh) return fail(SDM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    const ShardOffsets& so = h->res_offsets;
    if (so.count == 0) return fail(SDM_ERR_INVALID, "sdm_shard_resolve first");
    ShardTriOffsets to {};
    to.count = so.count;
    uint64_t tsum = 0;
    for (uint32_t s = 0; s < so.count; s++) { to.toff[s] = (uint32_t) tsum; tsum += triangle_counts[s]; }
    to.toff[so.count] = (uint32_t) tsum;
    const uint32_t total_v = so.voff[so.count], removed = so.poff[so.count];
    if (total_v > h->cap_uniq || tsum > h->cap_tris) return fail(SDM_ERR_CAPACITY, "call sdm_shard_reserve_welded with the totals first");
    cudaStream_t st = h->stream;
    CK(cudaEventRecord(h->ev0, st));
    const int b = h->welded_set, vb = b ^ 1;
    CK(cudaStreamWaitEvent(st, h->ev_copy_done[vb], 0));   // a download may still be reading the set the vertices move into
    if (removed) k_fix_scatter<<<h->g_light, 256, 0, st>>>(reinterpret_cast<const uint2*>(h->slot_ref.p), so, h->vidx.p);
    k_fix_vertices<<<h->g_light, 256, 0, st>>>(total_v, h->first_bits.p, h->first_prefix.p, h->out_pos[b].p, h->out_nrm[b].p, h->out_pos[vb].p, h->out_nrm[vb].p);
    k_fix_indices<<<h->g_light, 256, 0, st>>>(so, to, h->first_bits.p, h->first_prefix.p, h->vidx.p, h->out_idx[b].p);
    h->stats.kernel_launches += 3;
    CK(cudaEventRecord(h->ev1, st));
    CK(cudaEventRecord(h->ev_mesh_done[b], st));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->stats.last_gpu_ms = ms;
    h->welded_vset = vb;
    out_mesh->positions = h->out_pos[vb].p; out_mesh->normals = h->out_nrm[vb].p; out_mesh->indices = h->out_idx[b].p;
    out_mesh->vertex_count = total_v - removed; out_mesh->triangle_count = (uint32_t) tsum;
    out_mesh->on_device = 1; out_mesh->reserved = 4 | (vb << 1) | b;
    h->res_offsets.count = 0;
    return SDM_OK;
}

int sdm_shard_welded_buffers(SdmHandle* h, SdmShardBuffers* out) {
    if (!h || !out) return fail(SDM_ERR_INVALID, "null argument");
    out->positions = h->out_pos[h->welded_vset].p; out->normals = h->out_nrm[h->welded_vset].p; out->triangle_vertex_ids = h->out_idx[h->welded_set].p;
    out->capacity_vertices = h->cap_uniq; out->capacity_triangles = h->cap_tris;
    return SDM_OK;
}

int sdm_shard_reserve_welded(SdmHandle* h, uint32_t total_vertices, uint32_t total_triangles) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    if (total_vertices <= h->cap_uniq && total_triangles <= h->cap_tris) return SDM_OK;
    uint32_t want = h->cap_vox;
    while ((uint64_t) want * 2 < total_vertices || (uint64_t) want * 3 < total_triangles) {
        const uint32_t g = grown(want);
        if (g == want) return fail(SDM_ERR_CAPACITY, "merged mesh too large");
        want = g;
    }
    CK(cudaStreamSynchronize(h->stream));
    if (h->copy_stream) CK(cudaStreamSynchronize(h->copy_stream));
    // keep the local welded rows: detach them, re-allocate everything, copy back
    const int b = h->welded_set, vb = h->welded_vset;
    float* old_pos = h->out_pos[vb].p; float* old_nrm = h->out_nrm[vb].p; uint32_t* old_idx = h->out_idx[b].p;
    h->out_pos[vb].p = nullptr; h->out_pos[vb].n = 0; h->out_nrm[vb].p = nullptr; h->out_nrm[vb].n = 0; h->out_idx[b].p = nullptr; h->out_idx[b].n = 0;
    int rc = ensure_capacity(h, want);
    if (rc == SDM_OK) {
        cudaMemcpyAsync(h->out_pos[vb].p, old_pos, (size_t) h->welded_v * 12, cudaMemcpyDeviceToDevice, h->stream);
        cudaMemcpyAsync(h->out_nrm[vb].p, old_nrm, (size_t) h->welded_v * 12, cudaMemcpyDeviceToDevice, h->stream);
        cudaMemcpyAsync(h->out_idx[b].p, old_idx, (size_t) h->welded_t * 12, cudaMemcpyDeviceToDevice, h->stream);
        cudaStreamSynchronize(h->stream);
    }
    cudaFree(old_pos); cudaFree(old_nrm); cudaFree(old_idx);
    return rc;
}

int sdm_shard_weld(SdmHandle* h, uint32_t total_vertices, uint32_t total_triangles, SdmMesh* out_mesh) {
    if (!h || !out_mesh) return fail(SDM_ERR_INVALID, "null argument");
    if (total_vertices > h->cap_uniq || total_triangles > h->cap_tris) return fail(SDM_ERR_CAPACITY, "call sdm_shard_reserve first");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    CK(cudaEventRecord(h->ev0, s));
    h->host_range[3] = 0;
    uint32_t* counts = h->host_range;   // pinned scratch: {n_tris_raw, n_uniq}
    counts[0] = total_triangles; counts[1] = total_vertices;
    CK(cudaMemcpyAsync(&h->state.p->n_tris_raw, counts, 8, cudaMemcpyHostToDevice, s));   // n_tris_raw, n_uniq are adjacent
    int rc = enqueue_weld_clears(h, true);
    if (rc) return rc;
    k_first_slot_merged<<<h->g_light, 256, 0, s>>>(h->state.p, h->tri_uid.p, h->first_slot.p, h->tri_valid_bits.p, h->own_tris);
    h->stats.kernel_launches++;
    rc = enqueue_weld(h, false);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev1, s));
    uint32_t flags = 0;
    rc = fetch_state(h, &flags);
    if (rc) return rc;
    if (flags) return fail(SDM_ERR_CAPACITY, "weld table overflow");
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->stats.last_gpu_ms = ms;
    mesh_view(h, out_mesh);
    return SDM_OK;
}


// ---- ray-march viewer (CudaHandler::render, src/cuda/mod.rs:348-409) -----------------------------------------------------------
static_assert(sizeof(SdmRenderGlobals) == sizeof(RenderGlobals) && sizeof(SdmRenderCamera) == sizeof(RenderCamera), "render parameter layouts");
int sdm_render(SdmHandle* h, const SdmRenderGlobals* globals, const SdmRenderCamera* camera, unsigned char* out_rgba) {
    if (!h || !globals || !camera || !out_rgba) return fail(SDM_ERR_INVALID, "null argument");
    const uint32_t w = globals->render_texture_size[0], ht = globals->render_texture_size[1];
    if (w == 0 || ht == 0 || (w % 8u) || (ht % 16u) || (uint64_t) w * ht > (1ull << 28))
        return fail(SDM_ERR_INVALID, "image width must be a multiple of 8, height a multiple of 16 (the reference's 8 x 16 pixel blocks)");
    CK(cudaSetDevice(h->device));
    int rc = ensure_masks_any(h);
    if (rc) return rc;
    CK(h->render_target.reserve((size_t) w * ht));
    RenderGlobals g;
    RenderCamera c;
    memcpy(&g, globals, sizeof g);
    memcpy(&c, camera, sizeof c);
    NvtxRange nv(h, "render");
    // grid = (w * h) / BLOCK_SIZE blocks of BLOCK_SIZE threads (src/cuda/mod.rs:382-390)
    k_render<<<(w * ht) / 128u, 128, smem_for(h, 128), h->stream>>>(h->scene.p, h->render_target.p, g, c, w, ht, h->grid);
    h->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_rgba, h->render_target.p, (size_t) w * ht * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SDM_OK;
}

// ---- peer exchange -------------------------------------------------------------------------------------------------------------
int sdm_reserve(SdmHandle* h, uint32_t voxel_capacity) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    uint32_t want = 0;
    int rc = capacity_for(h, voxel_capacity, &want);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return ensure_capacity(h, want);
}

static size_t peer_rows_offset() { return (sizeof(PeerCtl) + 255) & ~(size_t) 255; }
static size_t peer_pairs_offset(uint32_t world, uint32_t cap_rows) { return peer_rows_offset() + (size_t) 2 * world * cap_rows * sizeof(uint4); }
static size_t peer_block_bytes(uint32_t world, uint32_t cap_rows) { return peer_pairs_offset(world, cap_rows) + (size_t) 2 * world * cap_rows * sizeof(uint2); }
static int peer_common_alloc(SdmHandle* h) {
    CK(h->peer_local.reserve(1));
    CK(h->peer_frac.reserve(2 * (SDM_PEER_MAX + 1)));
    CK(h->peer_t0.reserve(1));
    h->peer_last_epoch = 0;
    h->peer_rebalance = getenv("SDM_NO_REBALANCE") == nullptr;
    if (!h->host_peer_local) CK(cudaMallocHost(&h->host_peer_local, sizeof(PeerLocal)));
    return ensure_shard_scratch(h);
}

int sdm_peer_root_export(SdmHandle* h, uint32_t world, uint32_t cap_rows, SdmPeerExport* out) {
    if (!h || !out) return fail(SDM_ERR_INVALID, "null argument");
    if (world == 0 || world > SDM_PEER_MAX || cap_rows == 0) return fail(SDM_ERR_INVALID, "1..32 ranks, cap_rows > 0");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    // The resolve step (P2 of sdm_peer_step) keeps one word per key row in this handle's per-vertex scratch and the rows' keys in its
    // vertex table: world x cap_rows rows must fit (2 vertices per voxel of capacity; the table has 2 entries per vertex).  Growing here,
    // before the output buffers are exported, instead of failing at the first step.
    if ((uint64_t) world * cap_rows > h->cap_uniq) {
        uint32_t want = 0;
        int rc = capacity_for(h, ((uint64_t) world * cap_rows + 1) / 2, &want);
        if (rc) return rc;
        rc = ensure_capacity(h, want);
        if (rc) return rc;
    }
    const size_t bytes = peer_block_bytes(world, cap_rows);
    CK(h->peer_block.reserve(bytes));
    CK(cudaMemset(h->peer_block.p, 0, bytes));
    memset(out, 0, sizeof(*out));
    const void* ptrs[4] = { h->peer_block.p, h->out_pos[1].p, h->out_nrm[1].p, h->out_idx[1].p };
    for (int i = 0; i < 4; i++) {
        cudaIpcMemHandle_t mh;
        CK(cudaIpcGetMemHandle(&mh, const_cast<void*>(ptrs[i])));
        static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(out->handle[i], &mh, 64);
    }
    out->block_bytes = bytes; out->cap_vertices = h->cap_uniq; out->cap_triangles = h->cap_tris; out->cap_rows = cap_rows; out->world = world;
    return SDM_OK;
}

int sdm_peer_attach(SdmHandle* h, const SdmPeerExport* root, uint32_t rank, uint32_t world, SdmHandle* same_process_root) {
    if (!h || !root) return fail(SDM_ERR_INVALID, "null argument");
    if (world != root->world || rank >= world) return fail(SDM_ERR_INVALID, "rank / world do not match the export");
    CK(cudaSetDevice(h->device));
    int rc = peer_common_alloc(h);
    if (rc) return rc;
    void* base[4] = { nullptr, nullptr, nullptr, nullptr };
    if (rank == 0) {
        base[0] = h->peer_block.p; base[1] = h->out_pos[1].p; base[2] = h->out_nrm[1].p; base[3] = h->out_idx[1].p;
    } else if (same_process_root) {
        base[0] = same_process_root->peer_block.p; base[1] = same_process_root->out_pos[1].p; base[2] = same_process_root->out_nrm[1].p;
        base[3] = same_process_root->out_idx[1].p;
    } else {
        for (int i = 0; i < 4; i++) {
            if (h->ipc_mapped[i]) { cudaIpcCloseMemHandle(h->ipc_mapped[i]); h->ipc_mapped[i] = nullptr; }
            cudaIpcMemHandle_t mh;
            memcpy(&mh, root->handle[i], 64);
            CK(cudaIpcOpenMemHandle(&h->ipc_mapped[i], mh, cudaIpcMemLazyEnablePeerAccess));
            base[i] = h->ipc_mapped[i];
        }
    }
    if (!base[0]) return fail(SDM_ERR_STATE, "rank 0 has not exported yet");
    unsigned char* b0 = reinterpret_cast<unsigned char*>(base[0]);
    h->ctl = reinterpret_cast<PeerCtl*>(b0);
    h->peer_rows = reinterpret_cast<uint4*>(b0 + peer_rows_offset());
    h->peer_pairs = reinterpret_cast<uint2*>(b0 + peer_pairs_offset(world, root->cap_rows));
    h->root_pos = reinterpret_cast<float*>(base[1]); h->root_nrm = reinterpret_cast<float*>(base[2]); h->root_idx = reinterpret_cast<uint32_t*>(base[3]);
    h->peer_rank = rank; h->peer_world = world; h->peer_cap_rows = root->cap_rows; h->peer_cap_v = root->cap_vertices; h->peer_cap_t = root->cap_triangles;
    return SDM_OK;
}

int sdm_peer_detach(SdmHandle* h) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    for (void*& m : h->ipc_mapped) if (m) { cudaIpcCloseMemHandle(m); m = nullptr; }
    h->ctl = nullptr; h->peer_rows = nullptr; h->peer_pairs = nullptr; h->root_pos = h->root_nrm = nullptr; h->root_idx = nullptr;
    return SDM_OK;
}

int sdm_peer_step(SdmHandle* h, const SdmParams* params, uint32_t split_level, uint32_t epoch, int deliver, uint32_t phase_mask, int spin) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    if (!h->ctl) return fail(SDM_ERR_STATE, "sdm_peer_attach first");
    SdmParams p { SDM_MESH_GENERATION_BB_SIZE, SDM_MESH_GENERATION_INIT_FACTOR, 0 };
    if (params) p = *params;
    int rc = check_params(p);
    if (rc) return rc;
    if (split_level > p.levels) split_level = p.levels;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const uint32_t rank = h->peer_rank, world = h->peer_world, par = epoch & 1u, cap_rows = h->peer_cap_rows;
    PeerCtl* ctl = h->ctl;
    uint4* my_rows = h->peer_rows + ((size_t) par * world + rank) * cap_rows;
    const uint4* all_rows = h->peer_rows + (size_t) par * world * cap_rows;
    uint2* pairs = h->peer_pairs + (size_t) par * world * cap_rows;
    uint32_t* counter = h->shard_scratch.p + 3;
    h->peer_deliver = deliver;
    if (phase_mask & 1u) {   // P0: this rank's shard, welded locally; header to rank 0
        NvtxRange nv(h, "peer P0: shard + local weld");
        if ((uint64_t) p.init_factor * p.init_factor * p.init_factor > h->cap_vox) return fail(SDM_ERR_CAPACITY, "sdm_reserve first");
        prof_begin(h);
        CK(cudaEventRecord(h->ev0, s));
        k_peer_mark_start<<<1, 32, 0, s>>>(h->peer_t0.p);
        {   // shard fractions of this step from the compute times of the previous one (consecutive epochs only)
            float* fnew = h->peer_frac.p + (size_t) par * (SDM_PEER_MAX + 1);
            const float* fprev = h->peer_frac.p + (size_t) (par ^ 1u) * (SDM_PEER_MAX + 1);
            k_peer_rebalance<<<1, 32, 0, s>>>(ctl, par ^ 1u, world, fprev, fnew, (h->peer_rebalance && h->peer_last_epoch + 1 == epoch && epoch > 1) ? 1 : 0);
            h->shard_frac = fnew;
            h->peer_last_epoch = epoch;
            h->stats.kernel_launches += 2;
        }
        rc = reset_state(h);
        if (rc) return rc;
        mark(h, "start");
        rc = enqueue_init_field(h, p);
        if (rc) return rc;
        h->delta_override = h->slack_factor * ldexpf(p.bb_size / (float) p.init_factor, -(int) p.levels);
        for (uint32_t l = 0; l < split_level && !rc; l++) rc = enqueue_refine(h);
        if (rc) { h->delta_override = 0.0f; h->shard_frac = nullptr; return rc; }
        rc = enqueue_take_shard(h, rank, world);
        h->shard_frac = nullptr;
        if (rc) { h->delta_override = 0.0f; return rc; }
        for (uint32_t l = split_level; l < p.levels && !rc; l++) rc = enqueue_refine(h, l + 1 == p.levels);
        h->delta_override = 0.0f;
        if (rc) return rc;
        rc = enqueue_mesh_local(h, true);
        if (rc) return rc;
        h->out_sel = 0;   // set 0: the local weld; set 1: this rank's final rows (on rank 0 with deliver = 0: the merged mesh)
        rc = enqueue_weld(h, true);
        if (rc) return rc;
        k_shard_scratch_init<<<1, 32, 0, s>>>(h->shard_scratch.p);
        k_shard_xrange<<<h->g_light, 256, 0, s>>>(h->state.p, h->out_pos[0].p, h->shard_scratch.p);
        k_peer_publish_header<<<1, 32, 0, s>>>(h->state.p, h->shard_scratch.p, ctl, rank, par, h->peer_t0.p);
        k_peer_set_flag<<<1, 32, 0, s>>>(&ctl->flagA[rank], epoch);
        mark(h, "peer_header");
        h->stats.kernel_launches += 4;
    }
    if (phase_mask & 2u) {   // P1: key rows of the vertices inside another shard's x range -> this rank's slot on rank 0
        NvtxRange nv(h, "peer P1: interface key rows");
        if (spin) k_peer_wait<<<1, 32, 0, s>>>(ctl->flagA, world, epoch, h->state.p);
        k_peer_rows<<<h->g_light, 256, 0, s>>>(h->state.p, h->out_pos[0].p, ctl, rank, world, par, my_rows, cap_rows, counter);
        k_peer_publish_rows<<<1, 32, 0, s>>>(counter, ctl, rank, par, cap_rows);
        k_peer_set_flag<<<1, 32, 0, s>>>(&ctl->flagB[rank], epoch);
        mark(h, "peer_rows");
        h->stats.kernel_launches += 3 + (spin ? 1 : 0);
    }
    if ((phase_mask & 4u) && rank == 0) {   // P2: owners, removal bitmap, offsets, pairs
        NvtxRange nv(h, "peer P2: resolve");
        if (spin) k_peer_wait<<<1, 32, 0, s>>>(ctl->flagB, world, epoch, h->state.p);
        const uint32_t entries = std::min<uint32_t>(pow2_at_least(std::max<uint64_t>((uint64_t) world * cap_rows * 2, 1024)), h->table_entries);
        if ((uint64_t) entries * 4 < (uint64_t) world * cap_rows * 5 || (uint64_t) world * cap_rows > h->wref.n)
            return fail(SDM_ERR_CAPACITY, "rank 0's tables are too small for world x cap_rows key rows");
        k_peer_root_offsets<<<1, 32, 0, s>>>(ctl, world, par, h->cap_uniq, h->cap_tris);
        CK(dev_fill(s, h->first_bits.p, 0, ((size_t) h->cap_uniq / 32 + 2) * 4));
        k_peer_clear_table<<<h->g_light, 256, 0, s>>>(ctl, par, h->table2.p, entries);
        k_peer_res_insert<<<h->g_light, 256, 0, s>>>(ctl, world, par, cap_rows, all_rows, h->table2.p, entries, h->wref.p, &ctl->status[par]);
        k_peer_res_mark<<<h->g_light, 256, 0, s>>>(ctl, world, par, cap_rows, all_rows, h->table2.p, h->wref.p, h->first_bits.p);
        k_peer_scan_reset<<<1, 32, 0, s>>>(h->state.p, 2);
        k_bitscan<<<h->g_light, 256, 0, s>>>(h->state.p, h->first_bits.p, h->first_prefix.p, 2, next_epoch(h), h->tiles.p, &ctl->voff[par][world]);
        k_peer_root_goff<<<1, 32, 0, s>>>(ctl, world, par);
        k_peer_res_pairs<<<h->g_light, 256, 0, s>>>(ctl, world, par, cap_rows, all_rows, h->table2.p, h->wref.p, h->first_bits.p, h->first_prefix.p, pairs);
        if (deliver == 0) CK(cudaStreamWaitEvent(s, h->ev_copy_done[1], 0));   // the peers are about to overwrite the set a download may still read
        k_peer_set_flag<<<1, 32, 0, s>>>(&ctl->flagC, epoch);
        mark(h, "peer_resolve");
        h->stats.kernel_launches += 9 + (spin ? 1 : 0);
    }
    if (phase_mask & 8u) {   // P3: drop own duplicates, global indices, rows to their final place
        NvtxRange nv(h, "peer P3: apply + deliver");
        if (spin) k_peer_wait<<<1, 32, 0, s>>>(&ctl->flagC, 1, epoch, h->state.p);
        if (deliver != 0 || rank == 0) CK(cudaStreamWaitEvent(s, h->ev_copy_done[1], 0));
        CK(dev_fill(s, h->tri_valid_bits.p, 0, ((size_t) h->cap_uniq / 32 + 2) * 4));
        k_peer_apply_mark<<<h->g_light, 256, 0, s>>>(ctl, rank, par, pairs, h->tri_valid_bits.p, h->vidx.p, h->peer_local.p);
        k_peer_scan_reset<<<1, 32, 0, s>>>(h->state.p, 3);
        k_bitscan<<<h->g_light, 256, 0, s>>>(h->state.p, h->tri_valid_bits.p, h->tri_prefix.p, 3, next_epoch(h), h->tiles.p, &ctl->hdr[par][rank].V);
        // Rank 0 writes its rows where they belong (its second output set IS the merged mesh when deliver = 0); the other ranks compact
        // into their own second set and, for deliver = 0, push it to rank 0 in one coalesced copy.
        const bool direct = deliver == 0 && rank == 0;
        k_peer_apply_vertices<<<h->g_light * 2, 256, 0, s>>>(ctl, rank, par, h->tri_valid_bits.p, h->tri_prefix.p, h->out_pos[0].p, h->out_nrm[0].p, h->out_pos[1].p,
                                                             h->out_nrm[1].p, direct ? 1u : 0u);
        k_peer_apply_indices<<<h->g_light * 2, 256, 0, s>>>(ctl, rank, par, h->tri_valid_bits.p, h->tri_prefix.p, h->vidx.p, h->out_idx[0].p, h->out_idx[1].p,
                                                            direct ? 1u : 0u);
        mark(h, "peer_apply");
        if (deliver == 0 && rank != 0) {
            k_peer_push<<<h->num_sms * 4, 256, 0, s>>>(ctl, rank, par, reinterpret_cast<const uint32_t*>(h->out_pos[1].p), reinterpret_cast<const uint32_t*>(h->out_nrm[1].p),
                                                       h->out_idx[1].p, reinterpret_cast<uint32_t*>(h->root_pos), reinterpret_cast<uint32_t*>(h->root_nrm), h->root_idx);
            mark(h, "peer_push");
            h->stats.kernel_launches++;
        }
        k_peer_set_flag<<<1, 32, 0, s>>>(&ctl->flagD[rank], epoch);
        h->stats.kernel_launches += 6 + (spin ? 1 : 0);
    }
    if (phase_mask & 16u) {   // P4: rank 0 waits for everybody's rows; totals for the host
        if (spin && rank == 0) k_peer_wait<<<1, 32, 0, s>>>(ctl->flagD, world, epoch, h->state.p);
        k_peer_totals<<<1, 32, 0, s>>>(ctl, world, par, h->peer_local.p);
        CK(cudaEventRecord(h->ev1, s));
        CK(cudaEventRecord(h->ev_mesh_done[1], s));
        h->stats.kernel_launches += 1 + ((spin && rank == 0) ? 1 : 0);
    }
    CK(cudaGetLastError());
    return SDM_OK;
}

int sdm_peer_finish(SdmHandle* h, SdmPeerResult* out, SdmMesh* out_mesh) {
    if (!h || !out) return fail(SDM_ERR_INVALID, "null argument");
    if (!h->ctl) return fail(SDM_ERR_STATE, "sdm_peer_attach first");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->host_peer_local, h->peer_local.p, sizeof(PeerLocal), cudaMemcpyDeviceToHost, h->stream));
    uint32_t flags = 0;
    int rc = fetch_state(h, &flags);
    if (rc) return rc;
    const PeerLocal& L = *h->host_peer_local;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->stats.last_gpu_ms = ms;
    fill_stats(h, true);
    prof_end(h);
    memset(out, 0, sizeof(*out));
    out->status = L.status | flags; out->gpu_ms = ms;
    out->total_vertices = L.total_V; out->total_triangles = L.total_T;
    out->vertex_offset = L.goff; out->triangle_offset = L.toff; out->vertices = L.kept; out->triangles = L.T;
    if (out->status) {
        if ((out->status & 0xFFu) == ERR_HASH_FULL) h->table1_entries = h->table_entries;
        if ((out->status & 0xFFu) == ERR_LATTICE) { h->lattice_ok = false; h->lat_fail_bb = h->field_bb; h->lat_fail_init = h->field_init; }
        char buf[96];
        snprintf(buf, sizeof buf, "peer step failed on some rank: status 0x%x (capacity / table / non-finite vertices)", out->status);
        return fail(SDM_ERR_CAPACITY, buf);
    }
    adapt_table1(h);
    if (out_mesh) {
        const bool merged = h->peer_deliver == 0 && h->peer_rank == 0;
        out_mesh->positions = h->out_pos[1].p; out_mesh->normals = h->out_nrm[1].p; out_mesh->indices = h->out_idx[1].p;
        out_mesh->vertex_count = merged ? L.total_V : (h->peer_deliver ? L.kept : 0u);
        out_mesh->triangle_count = merged ? L.total_T : (h->peer_deliver ? L.T : 0u);
        out_mesh->on_device = 1; out_mesh->reserved = 1;
    }
    return SDM_OK;
}

int sdm_peer_download_async(SdmHandle* h, const SdmPeerResult* r, float* host_positions, float* host_normals, uint32_t* host_indices) {
    if (!h || !r) return fail(SDM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t cs = h->copy_stream;
    CK(cudaStreamWaitEvent(cs, h->ev_mesh_done[1], 0));
    if (r->vertices && host_positions) CK(cudaMemcpyAsync(host_positions + 3 * (size_t) r->vertex_offset, h->out_pos[1].p, (size_t) r->vertices * 12, cudaMemcpyDeviceToHost, cs));
    if (r->vertices && host_normals) CK(cudaMemcpyAsync(host_normals + 3 * (size_t) r->vertex_offset, h->out_nrm[1].p, (size_t) r->vertices * 12, cudaMemcpyDeviceToHost, cs));
    if (r->triangles && host_indices) CK(cudaMemcpyAsync(host_indices + 3 * (size_t) r->triangle_offset, h->out_idx[1].p, (size_t) r->triangles * 12, cudaMemcpyDeviceToHost, cs));
    CK(cudaEventRecord(h->ev_copy_done[1], cs));
    (void) cudaStreamQuery(cs);
    return SDM_OK;
}

// Debug / test access to intermediate device buffers of the last mesh stage (copies `bytes` bytes to host).
int sdm_debug_fetch(SdmHandle* h, const char* name, void* dst, size_t bytes) {
    if (!h || !name || !dst) return fail(SDM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    const std::string n(name);
    const void* src = nullptr;
    size_t have = 0;
    if (n == "ustart") { src = h->ustart.p; have = h->ustart.n * 4; }
    else if (n == "upos") { src = h->upos.p; have = h->upos.n * 4; }
    else if (n == "unrm") { src = h->unrm.p; have = h->unrm.n * 4; }
    else if (n == "tri_uid") { src = h->tri_uid.p; have = h->tri_uid.n * 4; }
    else if (n == "tri_off") { src = h->tri_off.p; have = h->tri_off.n * 4; }
    else if (n == "first_slot") { src = h->first_slot.p; have = h->first_slot.n * 4; }
    else if (n == "masks_fine") { src = h->masks_fine.p; have = h->masks_fine.n * 4; }
    else if (n == "masks_coarse") { src = h->masks_coarse.p; have = h->masks_coarse.n * 4; }
    else if (n == "state") { src = h->state.p; have = sizeof(DevState); }
    else if (n == "stragglers") { src = h->stragglers.p; have = h->stragglers.n * sizeof(Straggler); }
    else return fail(SDM_ERR_INVALID, "unknown buffer name");
    if (bytes > have) return fail(SDM_ERR_INVALID, "buffer smaller than requested");
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return SDM_OK;
}

// GPU self-test of the branch-free sqrt / division used by the culled fold: out[0] = sqrt mismatches over all 2^32 bit
// patterns (must be 0), out[1] = patterns sent to the slow path, out[2] = division mismatches over `div_samples` random
// (t, k) pairs (must be 0), out[3] = pairs sent to the slow path.
int sdm_selftest_math(SdmHandle* h, unsigned long long div_samples, unsigned long long* out4) {
    if (!h || !out4) return fail(SDM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    TempBuf<unsigned long long> tmp;
    CK(tmp.alloc(4));
    unsigned long long* d = tmp.p;
    CK(dev_fill(h->stream, d, 0, 32));
    k_selftest_math<<<h->num_sms * 8, 256, 0, h->stream>>>(d, div_samples);
    h->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out4, d, 32, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SDM_OK;
}

int sdm_set_profiling(SdmHandle* h, int enabled) {
    if (!h) return fail(SDM_ERR_INVALID, "null handle");
    h->profiling = enabled != 0;
    return SDM_OK;
}
// Per-kernel CUDA-event times (ms) of the last sdm_remesh, in launch order.  Returns the number of entries;
// names[i] points into storage owned by the handle (valid until the next remesh).
int sdm_get_kernel_times(SdmHandle* h, const char** names, float* ms, uint32_t capacity) {
    if (!h) return 0;
    const size_t n = std::min<size_t>(h->prof_ms.size(), capacity);
    for (size_t i = 0; i < n; i++) {
        if (names) names[i] = h->prof_names[i + 1].c_str();
        if (ms) ms[i] = h->prof_ms[i];
    }
    return (int) n;
}

int sdm_get_stats(SdmHandle* h, SdmStats* out) {
    if (!h || !out) return fail(SDM_ERR_INVALID, "null argument");
    *out = h->stats;
    return SDM_OK;
}

}  // extern "C"
