/*
 * sdfmesh.h - C ABI of libsdfmesh.so, the B200-native (sm_100a) drop-in for the mesh-generation hot
 * path of Meterius/bevy-signed-distance-mesh-generation.
 *
 * The reference has no host-side C ABI: its Rust `CudaHandler` (src/cuda/mod.rs) loads PTX with cudarc
 * and launches two `extern "C" __global__` kernels with by-value `#[repr(C)]` structs bindgen-generated
 * from cuda/includes/bindings.h.  This header therefore declares
 *   (1) the struct layouts of bindings.h, byte for byte (a Rust host can keep its generated types);
 *   (2) one C entry point per `CudaHandler` method on the path (what a Rust `extern "C"` block binds
 *       instead of cudarc launches) - see INTEGRATION.md for the Rust side;
 *   (3) the device-resident fast path (`sdm_remesh`) and the shard entry points used for multi-GPU.
 * The two reference kernel symbols themselves are also shipped with their original ABI in the compat
 * module (csrc/compat_module.cu -> compute_mesh_generation.ptx/.cubin), see INTEGRATION.md.
 *
 * All functions return 0 on success or a non-zero SdmStatus; sdm_last_error() gives the message.
 * Nothing here falls back to the CPU: without a CUDA device every compute entry point fails.
 * A handle is not thread-safe (the reference's handler lives in a Bevy NonSend resource,
 * src/renderer/mod.rs:230-235).  Calls are synchronous unless stated otherwise.
 */
#ifndef SDFMESH_H
#define SDFMESH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants re-exported by the reference's bindgen step (src/cuda/mod.rs:8) ------------------- */
#define SDM_BLOCK_SIZE 128                     /* bindings.h:7  BLOCK_SIZE */
#define SDM_MESH_GENERATION_INIT_FACTOR 32     /* bindings.h:9  MESH_GENERATION_INIT_FACTOR */
#define SDM_MESH_GENERATION_BB_SIZE 5.0f       /* bindings.h:10 MESH_GENERATION_BB_SIZE */

/* ---- FFI PODs: identical layout to bindings.h:43-64 (checked by static_asserts in csrc) ---------- */
typedef struct SdmPoint { float x, y, z; } SdmPoint;                        /* Point      12 B */
typedef struct SdmVoxelField {                                              /* VoxelField 32 B */
    SdmPoint voxel_size;        /* @0  */
    SdmPoint* voxels;           /* @16 min-corners of the active voxels */
    unsigned int voxel_count;   /* @24 */
} SdmVoxelField;
typedef struct SdmVertex { SdmPoint position; SdmPoint normal; } SdmVertex; /* Vertex     24 B */
typedef struct SdmTriangle { SdmVertex vertices[3]; } SdmTriangle;          /* Triangle   72 B */

typedef enum SdmStatus {
    SDM_OK = 0,
    SDM_ERR_CUDA = 1,          /* a CUDA runtime call failed (message has the cudaError string) */
    SDM_ERR_INVALID = 2,       /* bad argument */
    SDM_ERR_NO_DEVICE = 3,     /* no usable CUDA device: there is no CPU fallback */
    SDM_ERR_CAPACITY = 4,      /* a device buffer would exceed the configured limit */
    SDM_ERR_STATE = 5          /* call order violated (e.g. mesh before a field exists) */
} SdmStatus;

/* ---- scene description ----------------------------------------------------------------------------
 * The reference hard-codes one scene, sd_obj (cuda/modules/common.cu:222-226).  Here a scene is a
 * left fold over a primitive table, evaluated in index order:
 *     acc = FLT_MAX;  for i in 0..count:  acc = fold_i(acc, d_i(p))
 * with fold = min (signed_distance.cu:109 pattern) or smooth_min(acc, d, k) (signed_distance.cu:20-23).
 * Primitive distances are the reference's forms:
 *   SPHERE        length(p - a) - radius                      (common.cu:224, signed_distance.cu:82-84)
 *   BOX           sd_box(p, bp=a, bs=b)                       (signed_distance.cu:86-91)
 *   CAPSULE       sd_line(p, b0=a, b1=b) - radius             (signed_distance.cu:77-80, :109)
 *   BOX_SKELETON  sd_box_skeleton(p, bp=a, bs=b, lw=radius)   (signed_distance.cu:93-113, incl. its
 *                 `bs[(dir+1)%2]` indexing); must use fold=min unless it is primitive 0
 *   MANDELBULB    sd_mandelbulb(p / radius, 0) * radius       (signed_distance.cu:29-57; radius=0.4
 *                 gives sd_unit_mandelbulb)
 * sd_obj is { BOX_SKELETON(a=0, b=(3,1,.5), radius=.1, min), SPHERE(a=0, radius=1, smooth_min k=.5) }
 * and is what sdm_scene_default() returns / what a fresh handle uses.
 */
typedef enum SdmPrimKind {
    SDM_PRIM_SPHERE = 0, SDM_PRIM_BOX = 1, SDM_PRIM_CAPSULE = 2, SDM_PRIM_BOX_SKELETON = 3, SDM_PRIM_MANDELBULB = 4
} SdmPrimKind;
typedef enum SdmFoldOp { SDM_FOLD_MIN = 0, SDM_FOLD_SMOOTH_MIN = 1 } SdmFoldOp;

typedef struct SdmPrimitive {      /* 40 B */
    uint32_t kind;                 /* SdmPrimKind */
    uint32_t fold;                 /* SdmFoldOp */
    float k;                       /* smooth-min k (ignored for min) */
    float radius;
    float a[3];
    float b[3];
} SdmPrimitive;

/* Grid parameters; the reference compiles these in (bindings.h:9-10). */
typedef struct SdmParams {
    float bb_size;                 /* cube edge; level-0 grid spans [-bb/2, bb/2)^3; scenes of more than 24 primitives (culled fold): <= 32 */
    uint32_t init_factor;          /* level-0 voxels per axis */
    uint32_t levels;               /* refine steps performed by sdm_remesh */
} SdmParams;

/* Indexed mesh in the layout Bevy consumes (src/renderer/mod.rs:110-128): Float32x3 positions,
 * Float32x3 normals, u32 triangle-list indices, in the reference's weld order (src/cuda/mod.rs:263-296). */
typedef struct SdmMesh {
    float* positions;              /* vertex_count * 3 */
    float* normals;                /* vertex_count * 3 */
    uint32_t* indices;             /* triangle_count * 3 */
    uint32_t vertex_count;
    uint32_t triangle_count;
    int32_t on_device;             /* 1: pointers are device memory owned by the handle */
    int32_t reserved;              /* device meshes: which of the handle's two output sets holds it */
} SdmMesh;

typedef struct SdmHandle SdmHandle;

/* ---- lifecycle (CudaHandler::new, src/cuda/mod.rs:49-103) --------------------------------------- */
int sdm_create(int device_ordinal, SdmHandle** out_handle);
void sdm_destroy(SdmHandle* h);
const char* sdm_last_error(void);
const char* sdm_version(void);

/* ---- scene ------------------------------------------------------------------------------------- */
/* Writes the two primitives of the reference's sd_obj into out[0..1]; returns the count (2). */
uint32_t sdm_scene_default(SdmPrimitive* out, uint32_t capacity);
/* Compiles (segment set-up etc., with the reference's own operation order) and uploads the table. */
int sdm_set_scene(SdmHandle* h, const SdmPrimitive* prims, uint32_t count);
/* Scene SDF at arbitrary host points (n*3 floats in, n floats out) - used by the parity tests. */
int sdm_eval_sdf(SdmHandle* h, const float* points, uint32_t n, float* out_sd);
/* empirical_normal / closest_surface_point (signed_distance.cu:181-202, :227-240) at host points. */
int sdm_eval_normal(SdmHandle* h, const float* points, uint32_t n, float* out_normals);
int sdm_eval_project(SdmHandle* h, const float* points, uint32_t n, float* out_points, uint32_t* out_iters);

/* ---- the CudaHandler surface, host buffers in and out (drop-in semantics) ------------------------ */
/* create_cuda_voxel_field (src/cuda/mod.rs:105-122).  params==NULL -> the reference's 32 / 5.0.
 * The list is malloc'd by the library; release with sdm_voxel_field_free. */
int sdm_create_voxel_field(const SdmParams* params, SdmVoxelField* out_field);
void sdm_voxel_field_free(SdmVoxelField* field);
/* refine_voxel_field (src/cuda/mod.rs:124-202): the host list is replaced by the stably compacted
 * surviving children and voxel_size is halved.  Empty input is a no-op, as in the reference (:137). */
int sdm_refine_voxel_field(SdmHandle* h, SdmVoxelField* field);
/* voxel_field_to_mesh (src/cuda/mod.rs:204-346): welded indexed mesh in host memory owned by the
 * library (release with sdm_mesh_free).  Empty input gives an empty mesh (:327-345). */
int sdm_voxel_field_to_mesh(SdmHandle* h, const SdmVoxelField* field, SdmMesh* out_mesh);
void sdm_mesh_free(SdmMesh* mesh);

/* ---- device-resident path (nothing crosses PCIe between stages) ---------------------------------- */
/* Level-0 field on the device (same contents as sdm_create_voxel_field). */
int sdm_field_reset(SdmHandle* h, const SdmParams* params);
/* Upload a host list as the current device field. */
int sdm_field_upload(SdmHandle* h, const SdmVoxelField* field);
/* One subdivision level on the device field; *out_count (optional) receives the new voxel count. */
int sdm_field_refine(SdmHandle* h, uint32_t* out_count);
int sdm_field_count(SdmHandle* h, uint32_t* out_count, SdmPoint* out_voxel_size);
/* Copies the current active list to host (capacity in voxels). */
int sdm_field_download(SdmHandle* h, SdmPoint* out_voxels, uint32_t capacity);
/* Per-voxel marching-cubes case index (marching_cubes.cu:19-23) of the current field, to host. */
int sdm_field_cases(SdmHandle* h, uint8_t* out_cases, uint32_t capacity);
/* Mesh of the current device field.  out_mesh->on_device=1; buffers stay valid until the next
 * mesh/remesh call on this handle. */
int sdm_field_to_mesh(SdmHandle* h, SdmMesh* out_mesh);
/* Whole pipeline: level-0 field, params->levels refinements, mesh; device-resident result. */
int sdm_remesh(SdmHandle* h, const SdmParams* params, SdmMesh* out_mesh);
/* Copies a device-resident mesh into caller-provided host arrays (sizes from the SdmMesh counts). */
int sdm_mesh_download(SdmHandle* h, const SdmMesh* device_mesh, float* positions, float* normals, uint32_t* indices);
/* Streaming hand-off: the copies run on a second stream, ordered after the weld that produced `device_mesh`; mesh outputs
 * are double-buffered inside the handle, so the next sdm_remesh can be issued right away and overlaps the download.  At
 * most one download per output set may be in flight (i.e. call sdm_mesh_download_async once per remesh); host memory
 * should be pinned.  sdm_mesh_download_wait blocks until all issued downloads have landed. */
int sdm_mesh_download_async(SdmHandle* h, const SdmMesh* device_mesh, float* positions, float* normals, uint32_t* indices);
int sdm_mesh_download_wait(SdmHandle* h);
/* Consumer hand-off, file form: writes the mesh the way the reference does when its `Mesh` stage is advanced
 * (src/renderer/mod.rs:204 `obj.save("generated_mesh.obj")` on the ObjData built at src/cuda/mod.rs:303-326, obj crate 0.10.2):
 * all `v x y z` lines, the single `vt 0 0`, all `vn x y z` lines, `o default`, `g default`, then one
 * `f a/1/a b/1/b c/1/c` per triangle (1-based; position index == normal index, texture index 0 -> 1).  Numbers are written
 * as Rust's `{}` writes an f32: the shortest decimal digits that round-trip, positional notation, `NaN` / `inf` / `-inf`.
 * `mesh` may be a device mesh of this handle (it is downloaded first) or a host mesh (h may then be NULL).  An empty mesh
 * gives the header lines only (src/cuda/mod.rs:327-345). */
int sdm_mesh_save_obj(SdmHandle* h, const SdmMesh* mesh, const char* path);
/* FNV-1a-64 over raw bytes (the checksum of tests/golden and of bench.py's `mesh_fnv`). */
uint64_t sdm_hash_bytes(const void* data, size_t bytes);
/* The reference's raw output format: 5 Triangle slots per voxel, NaN-padded
 * (compute_mesh_generation.cu:64-120), to host (capacity in triangles, >= 5 * voxel_count). */
int sdm_field_triangle_soup(SdmHandle* h, SdmTriangle* out_triangles, uint32_t capacity);

/* ---- shards (multi-GPU: one process and one handle per GPU) -------------------------------------------------------
 * Every voxel is classified, projected and oriented from its own corner coordinates and the analytic SDF, so a
 * contiguous part of the active list can be meshed with no halo; only the weld (global first-occurrence vertex
 * order, src/cuda/mod.rs:263-296) needs all shards.  Flow per remesh:
 *   every rank : sdm_shard_remesh   - levels [0, split_level) redundantly, then its contiguous part of the
 *                                     level-`split_level` list refined to full depth and meshed up to (not incl.) the weld
 *   ranks > 0  : sdm_shard_prepare_send(vertex_offset) - make vertex ids global, mark dropped triangles
 *   transport  : positions / normals / triangle_vertex_ids (sdm_shard_buffers) are sent to rank 0 at the offsets
 *                given by the exclusive sums of the shards' counts (NCCL send/recv over NVLink; the Python host in
 *                bevy-signed-distance-mesh-generation_b200/parallel.py does this with torch.distributed)
 *   rank 0     : sdm_shard_reserve(totals) before receiving, sdm_shard_weld(totals) after: the merged mesh is
 *                byte-identical to the single-GPU mesh.
 */
typedef struct SdmShardInfo {
    uint32_t shard_index, shard_count, split_level;
    uint32_t split_total;            /* voxels in the level-`split_level` list */
    uint32_t voxel_begin, voxel_end; /* this shard's part of that list */
    uint32_t final_voxels;           /* this shard's voxels at the finest level */
    uint32_t unique_vertices;        /* rows of positions / normals */
    uint32_t raw_triangles;          /* rows of triangle_vertex_ids */
} SdmShardInfo;
typedef struct SdmShardBuffers {     /* device memory owned by the handle */
    float* positions;                /* [capacity_vertices][3] */
    float* normals;                  /* [capacity_vertices][3] */
    uint32_t* triangle_vertex_ids;   /* [capacity_triangles][3], post-flip order */
    uint32_t capacity_vertices, capacity_triangles;
} SdmShardBuffers;
int sdm_shard_remesh(SdmHandle* h, const SdmParams* params, uint32_t split_level, uint32_t shard_index, uint32_t shard_count,
                     SdmShardInfo* out_info);
int sdm_shard_buffers(SdmHandle* h, SdmShardBuffers* out);
int sdm_shard_prepare_send(SdmHandle* h, uint32_t vertex_offset);
/* Grows the handle's buffers to hold the merged mesh, preserving the local shard's rows. */
int sdm_shard_reserve(SdmHandle* h, uint32_t total_vertices, uint32_t total_triangles);
int sdm_shard_weld(SdmHandle* h, uint32_t total_vertices, uint32_t total_triangles, SdmMesh* out_mesh);

/* Distributed weld - the exchange of SURVEY.md section 8e; preferred, the calls above remain as its fallback.
 * Shards share vertices along their interfaces (and wherever two vertices have the same quantised key, src/cuda/mod.rs:270).
 * Global first-occurrence order is shard-major, so a key belongs to its copy in the LOWEST shard; the other copies are
 * removed and the indices that used them re-mapped:
 *   every rank : sdm_shard_local_weld      - welds its own shard (local order, local indices), reports the x range of its vertices
 *   rank 0     : sdm_shard_reserve_welded  - room for all shards' welded rows (call before sdm_shard_boundary_keys)
 *   every rank : sdm_shard_boundary_keys   - rows (kx, ky, kz, shard << 24 | local index) of its welded vertices whose x lies
 *                                            inside another shard's x range: only those can share a key with another shard
 *   transport  : the rows are gathered into rank 0's sdm_shard_key_scratch; the welded buffers (sdm_shard_welded_buffers) of
 *                shard s go to rank 0's welded buffers at the concatenated offsets (vertices / triangles of the shards
 *                before s) - still with local indices and duplicates, so this transfer needs no answer from rank 0 and
 *                overlaps the next call
 *   rank 0     : sdm_shard_resolve         - key table over the rows (smallest (shard, index) per key = the owner), removal
 *                                            bitmap over the concatenated vertex lists and its prefix pop-count
 *   rank 0     : sdm_shard_fixup           - drops the removed vertices (stable), makes every index global, hands out the mesh.
 * If a shard reports non-finite vertices or 2^24 or more vertices, the ranks fall back to sdm_shard_prepare_send /
 * sdm_shard_reserve / sdm_shard_weld.  Either way the merged mesh is byte-identical to the single-GPU mesh. */
typedef struct SdmShardWeld {
    uint32_t vertices;               /* welded vertices of this shard */
    uint32_t triangles;              /* kept triangles of this shard */
    uint32_t nonfinite;              /* welded vertices with a non-finite coordinate (forces the fallback) */
    float min_x, max_x;              /* range of x over the finite welded vertices (min_x > max_x: none) */
} SdmShardWeld;
int sdm_shard_local_weld(SdmHandle* h, SdmShardWeld* out);
int sdm_shard_reserve_welded(SdmHandle* h, uint32_t total_vertices, uint32_t total_triangles);
int sdm_shard_boundary_keys(SdmHandle* h, const float* lo, const float* hi, uint32_t interval_count /* <= 32 */,
                            uint32_t** out_rows_device /* [count][4] */, uint32_t* out_count);
int sdm_shard_key_scratch(SdmHandle* h, uint32_t rows, uint32_t** out_rows_device);
/* positions / normals: welded rows; triangle_vertex_ids: the welded index buffer */
int sdm_shard_welded_buffers(SdmHandle* h, SdmShardBuffers* out);
int sdm_shard_resolve(SdmHandle* h, const uint32_t* rows_device, uint32_t total_rows, const uint32_t* vertex_counts, uint32_t shard_count /* <= 32 */,
                      uint32_t* out_removed /* [shard_count] */);
int sdm_shard_fixup(SdmHandle* h, const uint32_t* triangle_counts /* [shard_count of the last sdm_shard_resolve] */, SdmMesh* out_mesh);

/* ---- ray-march viewer (CudaHandler::render, src/cuda/mod.rs:348-409; kernel compute_render, cuda/modules/compute_render.cu:21-97) ----
 * Sphere-traces the handle's current scene (<= 256 steps, depth limit 500, pixel-cone collision test), shades hits with the
 * finite-difference normal and the reference's two-colour ramp, ACES tone map, RGBA8.  The by-value structs have the layouts of
 * GlobalsBuffer / CameraBuffer (bindings.h:16-29).  The image is render_texture_size[0] x render_texture_size[1] (width a multiple
 * of 8, height a multiple of 16: the reference's 8 x 16 pixel blocks, common.cu:186-215); out_rgba = 4 * w * h host bytes.
 * The reference draws sd_obj plus the wire box of the meshing domain (compute_render.cu:3-19): that scene is the table
 * { BOX_SKELETON(0,(3,1,.5),.1) min, SPHERE(0,1) smooth_min .5, BOX_SKELETON(0,(5,5,5),.05) min }. */
typedef struct SdmRenderGlobals { unsigned long long tick; float time; unsigned int render_texture_size[2]; float render_screen_size[2]; } SdmRenderGlobals;
typedef struct SdmRenderCamera { float position[3]; float forward[3]; float up[3]; float right[3]; float fov; } SdmRenderCamera;
int sdm_render(SdmHandle* h, const SdmRenderGlobals* globals, const SdmRenderCamera* camera, unsigned char* out_rgba);

/* ---- peer exchange: the distributed weld without the host (one process per GPU on one node) ---------------------------------
 * The calls above route every count through the host.  Here rank 0 exports a control block, per-rank key-row slots and its second
 * output set (CUDA IPC); the other ranks map them, and one step runs from the first kernel to the last with device-side flags
 * only (csrc/sdm_kernels.cuh, "peer exchange"): local weld -> x ranges -> interface key rows to rank 0 -> rank 0 resolves owners and
 * global offsets -> every rank drops its own duplicates, makes its indices global and STORES its rows at their final offsets,
 * either straight into rank 0's output set over NVLink (deliver = 0) or into its own second output set, from where it copies its part
 * over its own PCIe link into a host buffer shared by the ranks (deliver = 1, sdm_peer_download_async).  The result is byte-identical
 * to the single-GPU mesh.  Capacities are fixed at export time (sdm_reserve first); a step that does not fit fails on all ranks. */
typedef struct SdmPeerExport {
    unsigned char handle[4][64];     /* cudaIpcMemHandle_t of: control block, positions, normals, indices of rank 0's output set */
    uint64_t block_bytes;
    uint32_t cap_vertices, cap_triangles, cap_rows, world;
} SdmPeerExport;
typedef struct SdmPeerResult {
    uint32_t status;                 /* 0 = ok; otherwise the OR of the ranks' error bits (the step produced nothing) */
    uint32_t total_vertices, total_triangles;   /* the merged mesh */
    uint32_t vertex_offset, triangle_offset;    /* where this rank's rows start in it */
    uint32_t vertices, triangles;               /* this rank's rows (duplicates owned by lower ranks removed) */
    float gpu_ms;                               /* CUDA-event time of the step on this rank's stream */
} SdmPeerResult;
int sdm_reserve(SdmHandle* h, uint32_t voxel_capacity);   /* grow the handle's buffers now (2 vertices / 3 triangles per voxel of capacity) */
int sdm_peer_root_export(SdmHandle* h, uint32_t world, uint32_t cap_rows_per_rank, SdmPeerExport* out);            /* rank 0 */
/* every rank (rank 0 too).  same_process_root != NULL: ranks emulated by several handles of one process (tests): rank 0's pointers are
 * used directly, and the caller must issue the phases of a step one by one (phase_mask) with a synchronisation in between. */
int sdm_peer_attach(SdmHandle* h, const SdmPeerExport* root, uint32_t rank, uint32_t world, SdmHandle* same_process_root);
/* Unmaps rank 0's memory (before rank 0 grows its buffers and exports again after a step that did not fit). */
int sdm_peer_detach(SdmHandle* h);
/* phase_mask: bit p = enqueue phase p (0..4); spin != 0: each phase first waits on the device for the flags it depends on
 * (one rank per GPU only).  epoch: 1, 2, 3, ... the same on all ranks. */
int sdm_peer_step(SdmHandle* h, const SdmParams* params, uint32_t split_level, uint32_t epoch, int deliver, uint32_t phase_mask, int spin);
int sdm_peer_finish(SdmHandle* h, SdmPeerResult* out, SdmMesh* out_mesh /* rank 0 with deliver = 0: the merged mesh; else this rank's rows */);
/* deliver = 1: this rank's rows into the shared host arrays at its offsets (asynchronous; sdm_mesh_download_wait) */
int sdm_peer_download_async(SdmHandle* h, const SdmPeerResult* r, float* host_positions, float* host_normals, uint32_t* host_indices);

/* ---- counters for the bench ------------------------------------------------------------------------ */
typedef struct SdmStats {
    uint64_t kernel_launches;      /* kernels launched by this handle since creation */
    uint64_t sdf_evals;            /* SDF evaluations of the last remesh/mesh as the algorithm states them per distinct vertex / triangle
                                      (27 per parent, 13 per Newton step, 12 per normal, 12 per triangle - of which the orientation
                                      test usually needs 6: prim_evals holds what was really folded) */
    uint32_t level_counts[16];     /* active voxels after each level of the last remesh (index 0 = level 0) */
    uint32_t unique_vertices;      /* distinct edge midpoints projected in the last mesh */
    uint32_t raw_triangles;        /* triangles before the finite-vertex filter */
    float last_gpu_ms;             /* CUDA-event time of the last remesh / mesh call on the handle's stream */
    uint32_t escaped_vertices;     /* vertices whose Newton iterate left the region of their inherited primitive list (general path) */
    uint64_t prim_evals[6];        /* (primitive, point) distance evaluations actually folded by the last remesh, per stage:
                                      refine, classify, project, project tail, vertex normals, orient (after culling) */
    uint32_t list_fallback_tiles;  /* warp tiles of the mesh stage that used the cell masks although list records existed */
    uint32_t stragglers;           /* vertices handed to the Newton tail kernel (>= 40 iterations, or escaped) */
    uint64_t newton_iterations;    /* closest_surface_point iterations of the last mesh (after exact cycle short-cuts) */
} SdmStats;
int sdm_get_stats(SdmHandle* h, SdmStats* out);
/* Per-kernel timing for the bench's roofline line: when enabled, sdm_remesh records a CUDA event on the handle's
 * stream after every kernel it enqueues; sdm_get_kernel_times returns (name, ms) per kernel of the last remesh. */
int sdm_set_profiling(SdmHandle* h, int enabled);
/* GPU self-test of the branch-free IEEE sqrt / division of the culled fold (csrc/sdm_device.cuh): out4 = {sqrt
 * mismatches over all 2^32 bit patterns, sqrt patterns sent to the slow path, division mismatches over div_samples
 * random pairs, pairs sent to the slow path}.  Mismatch counts must be 0. */
int sdm_selftest_math(SdmHandle* h, unsigned long long div_samples, unsigned long long* out4);
/* Test access to intermediate device buffers of the last mesh stage ("ustart", "upos", "unrm", "tri_uid", "tri_off",
 * "first_slot"): copies `bytes` bytes to host memory. */
int sdm_debug_fetch(SdmHandle* h, const char* name, void* dst, size_t bytes);
int sdm_get_kernel_times(SdmHandle* h, const char** names, float* ms, uint32_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* SDFMESH_H */
