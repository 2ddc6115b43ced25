// compat_module.cu - the reference's kernel-level ABI, re-implemented for sm_100a.
//
// The reference's Rust host does not call a C function: it loads the PTX module "compute_mesh_generation" with cudarc
// and launches two `extern "C" __global__` symbols by name with by-value #[repr(C)] structs
// (src/cuda/mod.rs:68-90, 149-177, 226-250; struct layouts cuda/includes/bindings.h:43-64).  This translation unit
// exports exactly those two symbols with exactly those parameter layouts, launch shape (grid = ceil(n / 128), block =
// 128, no dynamic shared memory) and output conventions, so the UNMODIFIED Rust host runs on a B200 by swapping
// assets/cuda/compiled/compute_mesh_generation.ptx for the file built from this source (csrc/Makefile: `make compat`).
//
//   compute_refine_voxel_field_by_sdf(VoxelField in, VoxelField out)
//       out.voxels[id*8 + i*4 + j*2 + k] = child min-corner, or (INF, INF, INF) if its 8 corners agree in sign
//       (compute_mesh_generation.cu:12-62); writes are bounded by out.voxel_count (:53)
//   compute_surface_triangles_from_voxel_field_by_sdf(VoxelField field, Triangle* triangles)
//       5 Triangle slots per voxel; unused slots = { POINT_NAN, POINT_NAN } = vertex 0 NaN, vertices 1..2 zero (:116-118)
//
// The scene is the reference's compiled-in sd_obj (common.cu:222-226), as in the reference.  Inside, the kernels use the
// same device code as libsdfmesh (27-point lattice instead of 64 corner evaluations, one projection per distinct edge of
// a voxel instead of one per triangle corner); results are bit-identical to the reference kernels compiled with IEEE
// arithmetic (tests/test_gpu_compat.py loads the module through the CUDA driver API, as cudarc does).
#include "sdm_device.cuh"

using namespace sdm;

extern "C" {
struct Point { float x, y, z; };
struct VoxelField { Point voxel_size; Point* voxels; unsigned int voxel_count; };
struct Vertex { Point position; Point normal; };
struct Triangle { Vertex vertices[3]; };
}
static_assert(sizeof(VoxelField) == 32 && sizeof(Triangle) == 72, "bindings.h layouts");

__constant__ unsigned long long k_mc_packed[256] = {
#define SDM_TABLE_AS_LIST
#include "mc_tables_list.inc"
};

namespace {

struct SdObjScene { SceneHeader hdr; DevRun runs[2]; DevPrim prims[13]; };   // 16 + 32 + 832 bytes

__device__ __forceinline__ void put_capsule(DevPrim& d, float ax, float ay, float az, float bx, float by, float bz, float lw) {
    const float ex = bx - ax, ey = by - ay, ez = bz - az;
    const float len = sqrtf(ex * ex + ey * ey + ez * ez);          // length(b1 - b0)          (signed_distance.cu:78)
    d.v0[0] = ax; d.v0[1] = ay; d.v0[2] = az;
    d.v1[0] = ex / len; d.v1[1] = ey / len; d.v1[2] = ez / len;    // (b1 - b0) / len          (:79)
    d.v2[0] = 0.f; d.v2[1] = 0.f; d.v2[2] = 0.f;
    d.s0 = lw; d.s1 = len; d.k = 0.0f;
    d.kind = SDM_PRIM_CAPSULE; d.fold = SDM_FOLD_MIN; d.pad0 = 0; d.pad1 = 0;
}

// sd_obj = smooth_min(sd_box_skeleton(p, 0, (3, 1, .5), .1), length(p) - 1, .5) as a 13-primitive table in shared memory
__device__ __forceinline__ SceneView build_sd_obj(SdObjScene* s) {
    const int t = threadIdx.x;
    if (t < 12) {
        const float bs[3] = { 3.0f, 1.0f, 0.5f };
        const float bpl[3] = { 0.0f - bs[0] / 2.0f, 0.0f - bs[1] / 2.0f, 0.0f - bs[2] / 2.0f };   // bp - bs / 2.0f (:94)
        const int dir = t >> 2, c0 = (t >> 1) & 1, c1 = t & 1;                                     // loop order of :97-99
        float m0[3] = { bpl[0], bpl[1], bpl[2] };
        m0[(dir + 1) % 3] += c0 ? bs[(dir + 1) % 2] : 0.0f;                                        // sic: % 2 (:101)
        m0[(dir + 2) % 3] += c1 ? bs[(dir + 2) % 3] : 0.0f;
        float m1[3] = { m0[0], m0[1], m0[2] };
        m1[dir] += bs[dir];
        put_capsule(s->prims[t], m0[0], m0[1], m0[2], m1[0], m1[1], m1[2], 0.1f);
    } else if (t == 12) {
        DevPrim& d = s->prims[12];
        d.v0[0] = d.v0[1] = d.v0[2] = 0.0f; d.s0 = 1.0f;
        d.v1[0] = d.v1[1] = d.v1[2] = 0.0f; d.s1 = 0.0f;
        d.v2[0] = d.v2[1] = d.v2[2] = 0.0f; d.k = 0.5f;
        d.kind = SDM_PRIM_SPHERE; d.fold = SDM_FOLD_SMOOTH_MIN; d.pad0 = 0; d.pad1 = 0;
    } else if (t == 13) {
        s->runs[0] = DevRun { SDM_PRIM_CAPSULE, SDM_FOLD_MIN, 0u, 12u | ((uint32_t) RUN_SHARED_RADIUS_MIN << 24) };
        s->runs[1] = DevRun { SDM_PRIM_SPHERE, SDM_FOLD_SMOOTH_MIN, 12u, 1u };
    }
    __syncthreads();
    SceneView v;
    v.prims = s->prims; v.runs = s->runs; v.nruns = 2; v.nprims = 13;
    v.wmask = nullptr; v.W = 0; v.tlist = nullptr; v.tcount = nullptr;
    return v;
}

}  // namespace

extern "C" __global__ void __launch_bounds__(128) compute_refine_voxel_field_by_sdf(const VoxelField input_field, VoxelField output_field) {
    __shared__ SdObjScene scene;
    const SceneView sc = build_sd_obj(&scene);
    const unsigned int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= input_field.voxel_count) return;
    const float osx = input_field.voxel_size.x / 2.0f, osy = input_field.voxel_size.y / 2.0f, osz = input_field.voxel_size.z / 2.0f;   // :20
    const Point base = input_field.voxels[id];
    uint32_t m27 = 0;
#pragma unroll 1
    for (int a = 0; a < 3; a++) {
        float px[9], py[9], pz[9], f[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            px[q] = base.x + (float) a * osx; py[q] = base.y + (float) (q / 3) * osy; pz[q] = base.z + (float) (q % 3) * osz;
        }
        eval_scene<9>(sc, px, py, pz, f);
#pragma unroll
        for (int q = 0; q < 9; q++) m27 |= (uint32_t) (f[q] <= 0.0f) << (a * 9 + q);
    }
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int ch = 0; ch < 8; ch++) {
        const uint32_t M = child_mask(ch >> 2, (ch >> 1) & 1, ch & 1);
        const uint32_t sgn = m27 & M;
        const bool border = sgn != 0u && sgn != M;
        const unsigned int n_id = id * 8u + (unsigned int) ch;   // id*8 + i*4 + j*2 + k (:51)
        if (n_id < output_field.voxel_count) {
            Point o;
            o.x = border ? base.x + (float) (ch >> 2) * osx : inf;
            o.y = border ? base.y + (float) ((ch >> 1) & 1) * osy : inf;
            o.z = border ? base.z + (float) (ch & 1) * osz : inf;
            output_field.voxels[n_id] = o;
        }
    }
}

extern "C" __global__ void __launch_bounds__(128) compute_surface_triangles_from_voxel_field_by_sdf(VoxelField field, Triangle* triangles) {
    __shared__ SdObjScene scene;
    const SceneView sc = build_sd_obj(&scene);
    const unsigned int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= field.voxel_count) return;
    const Point base = field.voxels[id];
    const float sx = field.voxel_size.x, sy = field.voxel_size.y, sz = field.voxel_size.z;
    float cx[8], cy[8], cz[8], f[8];
#pragma unroll
    for (int c = 0; c < 8; c++) voxel_corner(base.x, base.y, base.z, sx, sy, sz, c, cx[c], cy[c], cz[c]);
    eval_scene<8>(sc, cx, cy, cz, f);
    uint32_t cube_index = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) cube_index |= (uint32_t) (f[c] <= 0.0f) << c;
    const unsigned long long packed = k_mc_packed[cube_index];
    uint32_t ntri = 0;
    while (ntri < 5 && ((packed >> (12 * ntri)) & 0xFFFull) != 0ull) ntri++;   // a triangle never has three edge-0 corners
    Triangle* out = triangles + 5u * id;
    const float qnan = __int_as_float(0x7fc00000);
    for (uint32_t t = 0; t < 5; t++) {
        if (t >= ntri) {
            Triangle pad;
            pad.vertices[0].position = Point { qnan, qnan, qnan }; pad.vertices[0].normal = Point { qnan, qnan, qnan };
            pad.vertices[1].position = Point { 0.f, 0.f, 0.f }; pad.vertices[1].normal = Point { 0.f, 0.f, 0.f };
            pad.vertices[2] = pad.vertices[1];
            out[t] = pad;
            continue;
        }
        float v[3][3], n[3][3];
        for (int j = 0; j < 3; j++) {
            const int e = (int) ((packed >> (12 * t + 4 * j)) & 0xFull);
            const int c0 = (e < 4) ? ((e == 3) ? 0 : e) : (e < 8 ? ((e == 7) ? 4 : e) : e - 8);
            const int c1 = (e < 4) ? ((e == 3) ? 3 : e + 1) : (e < 8 ? ((e == 7) ? 7 : e + 1) : e - 4);
            float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (c == c0) { ax = cx[c]; ay = cy[c]; az = cz[c]; }
                if (c == c1) { bx = cx[c]; by = cy[c]; bz = cz[c]; }
            }
            float gx = ax * (1.0f - 0.5f) + bx * 0.5f, gy = ay * (1.0f - 0.5f) + by * 0.5f, gz = az * (1.0f - 0.5f) + bz * 0.5f;
            bool collision = false;
            for (int i = 0; !collision && i < 10000; i++) collision = newton_step(sc, gx, gy, gz);   // closest_surface_point
            v[j][0] = gx; v[j][1] = gy; v[j][2] = gz;
            empirical_normal(sc, gx, gy, gz, n[j][0], n[j][1], n[j][2]);
        }
        const float ax = v[1][0] - v[0][0], ay = v[1][1] - v[0][1], az = v[1][2] - v[0][2];
        const float bx = v[2][0] - v[0][0], by = v[2][1] - v[0][1], bz = v[2][2] - v[0][2];
        const float kx = ay * bz - by * az, ky = az * bx - bz * ax, kz = ax * by - bx * ay;
        const float inv = 1.0f / sqrtf(dot3(kx, ky, kz, kx, ky, kz));
        float nx, ny, nz;
        empirical_normal(sc, (v[0][0] + v[1][0] + v[2][0]) / 3.0f, (v[0][1] + v[1][1] + v[2][1]) / 3.0f, (v[0][2] + v[1][2] + v[2][2]) / 3.0f, nx, ny, nz);
        const bool flip = dot3(kx * inv, ky * inv, kz * inv, nx, ny, nz) <= 0.0f;
        const int a = flip ? 2 : 0, c = flip ? 0 : 2;
        Triangle tri;
        tri.vertices[0].position = Point { v[a][0], v[a][1], v[a][2] }; tri.vertices[0].normal = Point { n[a][0], n[a][1], n[a][2] };
        tri.vertices[1].position = Point { v[1][0], v[1][1], v[1][2] }; tri.vertices[1].normal = Point { n[1][0], n[1][1], n[1][2] };
        tri.vertices[2].position = Point { v[c][0], v[c][1], v[c][2] }; tri.vertices[2].normal = Point { n[c][0], n[c][1], n[c][2] };
        out[t] = tri;
    }
}
