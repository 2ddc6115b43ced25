// TEST INFRASTRUCTURE ONLY (oracle/_ref): the reference's own mesh-generation translation unit compiled by nvcc for
// sm_100a, UNMODIFIED and BY PATH, behind a small host launcher with a C ABI - the GPU-side oracle.
//
// Flags (oracle/Makefile): -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, i.e. IEEE arithmetic without
// contraction.  The reference's own build uses --use_fast_math (build.rs:113,118), whose results are not
// reproducible across compilers/architectures; the parity target of this repo is the IEEE evaluation of the same
// source, which (a) equals the host-compiled reference bit for bit on scenes that only use + - * / sqrt
// (tests/test_gpu_ref_oracle.py checks this on the B200) and (b) is the only meaningful oracle for the Mandelbulb,
// whose acos/atan2/pow/sin/cos/log come from CUDA's libdevice on the GPU and from libm on the CPU.
//
// Part 1 launches the two reference kernels exactly as src/cuda/mod.rs does (grid = ceil(n/128), block = 128).
// Part 2 re-states the two kernel bodies (compute_mesh_generation.cu:12-62, 64-120) as templates over the SDF
// functor, built only from the reference's own device functions (march_cube, closest_surface_point,
// empirical_normal, sd_*), because the reference hard-wires sd_obj; instantiated with sd_obj the templates must
// reproduce part 1 byte for byte (same test), which validates them for the other scene (sd_unit_mandelbulb).
#include REF_MESH_CU

#include <cstdio>
#include <cuda_runtime.h>

namespace {

struct SdObj { __device__ float operator()(const vec3 p) const { return sd_obj(p); } };
struct SdUnitMandelbulb { __device__ float operator()(const vec3 p) const { return sd_unit_mandelbulb(p); } };

template <class Sd> __global__ void tpl_refine(const VoxelField in, VoxelField out, Sd sd) {
    const vec3 child_size = from_point(in.voxel_size) / 2.0f;
    const unsigned int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= in.voxel_count) return;
    const vec3 base = from_point(in.voxels[id]);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                const vec3 lo = base + vec3 { i, j, k } * child_size;
                const vec3 hi = base + vec3 { i + 1, j + 1, k + 1 } * child_size;
                const bool first = sd(lo) <= 0.0f;
                bool border = false;
                for (int c = 1; c < 8 && !border; c++)
                    border = first != (sd(vec3 { c & 1 ? hi[0] : lo[0], c & 2 ? hi[1] : lo[1], c & 4 ? hi[2] : lo[2] }) <= 0.0f);
                const unsigned int slot = id * 8 + i * 4 + j * 2 + k;
                if (slot < out.voxel_count)
                    out.voxels[slot] = { border ? lo.x : INFINITY, border ? lo.y : INFINITY, border ? lo.z : INFINITY };
            }
}

template <class Sd> __global__ void tpl_mesh(VoxelField field, Triangle* triangles, Sd sd) {
    const vec3 size = from_point(field.voxel_size);
    const unsigned int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= field.voxel_count) return;
    const vec3 base = from_point(field.voxels[id]);
    McCube cube;
    for (int c = 0; c < 8; c++) {
        vec3 v = base;
        v[0] += (c % 4) == 1 || (c % 4) == 2 ? size.x : 0.0f;
        v[1] += (c % 4) >= 2 ? size.y : 0.0f;
        v[2] += c >= 4 ? size.z : 0.0f;
        cube.vertices[c] = v;
        cube.values[c] = sd(v);
    }
    Triangle* mine = triangles + 5 * id;
    const unsigned int count = march_cube(cube, mine);
    for (unsigned int t = 0; t < count; t++) {
        vec3 v0 = closest_surface_point(sd, from_point(mine[t].vertices[0].position));
        vec3 v1 = closest_surface_point(sd, from_point(mine[t].vertices[1].position));
        vec3 v2 = closest_surface_point(sd, from_point(mine[t].vertices[2].position));
        vec3 n0 = empirical_normal(sd, v0), n1 = empirical_normal(sd, v1), n2 = empirical_normal(sd, v2);
        const vec3 face = normalize(cross(v1 - v0, v2 - v0));
        const vec3 field_normal = empirical_normal(sd, (v0 + v1 + v2) / 3.0f);
        const bool flip = dot(face, field_normal) <= 0.0f;
        mine[t].vertices[0] = { to_point(flip ? v2 : v0), to_point(flip ? n2 : n0) };
        mine[t].vertices[1] = { to_point(v1), to_point(n1) };
        mine[t].vertices[2] = { to_point(flip ? v0 : v2), to_point(flip ? n0 : n2) };
    }
    for (unsigned int t = count; t < 5; t++) mine[t] = { POINT_NAN, POINT_NAN };
}

template <class Sd> __global__ void tpl_sdf(const Point* pts, unsigned n, float* out, Sd sd) {
    const unsigned int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id < n) out[id] = sd(from_point(pts[id]));
}

#define RCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "ref_gpu: %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

unsigned grid_for(unsigned n) { return (n + BLOCK_SIZE - 1) / BLOCK_SIZE; }

}  // namespace

extern "C" {

// scene: 0 = the unmodified reference kernels (sd_obj); 1 = template<sd_obj>; 2 = template<sd_unit_mandelbulb>
int refgpu_refine(int scene, const float* voxels, unsigned n, const float* voxel_size, float* out_voxels /* 8n*3 */) {
    if (n == 0) return 0;
    Point *d_in = nullptr, *d_out = nullptr;
    RCK(cudaMalloc(&d_in, (size_t) n * 12));
    RCK(cudaMalloc(&d_out, (size_t) n * 8 * 12));
    RCK(cudaMemcpy(d_in, voxels, (size_t) n * 12, cudaMemcpyHostToDevice));
    VoxelField in { { voxel_size[0], voxel_size[1], voxel_size[2] }, d_in, n };
    VoxelField out { { 0.0f, 0.0f, 0.0f }, d_out, n * 8u };
    if (scene == 0) compute_refine_voxel_field_by_sdf<<<grid_for(n), BLOCK_SIZE>>>(in, out);
    else if (scene == 1) tpl_refine<<<grid_for(n), BLOCK_SIZE>>>(in, out, SdObj());
    else tpl_refine<<<grid_for(n), BLOCK_SIZE>>>(in, out, SdUnitMandelbulb());
    RCK(cudaGetLastError());
    RCK(cudaMemcpy(out_voxels, d_out, (size_t) n * 8 * 12, cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_out);
    return 0;
}

int refgpu_mesh(int scene, const float* voxels, unsigned n, const float* voxel_size, float* out_triangles /* 5n*18 */) {
    if (n == 0) return 0;
    Point* d_in = nullptr;
    Triangle* d_tri = nullptr;
    RCK(cudaMalloc(&d_in, (size_t) n * 12));
    RCK(cudaMalloc(&d_tri, (size_t) n * 5 * sizeof(Triangle)));
    RCK(cudaMemcpy(d_in, voxels, (size_t) n * 12, cudaMemcpyHostToDevice));
    VoxelField in { { voxel_size[0], voxel_size[1], voxel_size[2] }, d_in, n };
    if (scene == 0) compute_surface_triangles_from_voxel_field_by_sdf<<<grid_for(n), BLOCK_SIZE>>>(in, d_tri);
    else if (scene == 1) tpl_mesh<<<grid_for(n), BLOCK_SIZE>>>(in, d_tri, SdObj());
    else tpl_mesh<<<grid_for(n), BLOCK_SIZE>>>(in, d_tri, SdUnitMandelbulb());
    RCK(cudaGetLastError());
    RCK(cudaMemcpy(out_triangles, d_tri, (size_t) n * 5 * sizeof(Triangle), cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_tri);
    return 0;
}

int refgpu_sdf(int scene, const float* pts, unsigned n, float* out) {
    if (n == 0) return 0;
    Point* d_in = nullptr;
    float* d_out = nullptr;
    RCK(cudaMalloc(&d_in, (size_t) n * 12));
    RCK(cudaMalloc(&d_out, (size_t) n * 4));
    RCK(cudaMemcpy(d_in, pts, (size_t) n * 12, cudaMemcpyHostToDevice));
    if (scene == 2) tpl_sdf<<<grid_for(n), BLOCK_SIZE>>>(d_in, n, d_out, SdUnitMandelbulb());
    else tpl_sdf<<<grid_for(n), BLOCK_SIZE>>>(d_in, n, d_out, SdObj());
    RCK(cudaGetLastError());
    RCK(cudaMemcpy(out, d_out, (size_t) n * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_out);
    return 0;
}

}  // extern "C"
