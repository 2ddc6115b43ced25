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
// Part 2 (ref_functor.inc) re-states the two kernel bodies (compute_mesh_generation.cu:12-62, 64-120) as templates
// over the SDF functor, built only from the reference's own device functions (march_cube, closest_surface_point,
// empirical_normal, sd_*), because the reference hard-wires sd_obj; instantiated with sd_obj the templates must
// reproduce part 1 byte for byte (same test), which validates them for the other scenes: sd_unit_mandelbulb and
// the primitive-table fold (SdTable: the 1024-primitive scene of BASELINE configs[2], evaluated UN-CULLED).
#include REF_MESH_CU

#include <cstdio>
#include <cuda_runtime.h>

#include "ref_functor.inc"

namespace {

#define RCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "ref_gpu: %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

unsigned grid_for(unsigned n) { return (n + BLOCK_SIZE - 1) / BLOCK_SIZE; }

}  // namespace

extern "C" {

// primitive table for scene 3 (device copy owned by this library)
static SdmPrimitive* g_table = nullptr;
static unsigned g_table_count = 0;
int refgpu_set_table(const SdmPrimitive* prims, unsigned count) {
    if (g_table) { cudaFree(g_table); g_table = nullptr; }
    g_table_count = count;
    if (count == 0) return 0;
    RCK(cudaMalloc(&g_table, (size_t) count * sizeof(SdmPrimitive)));
    RCK(cudaMemcpy(g_table, prims, (size_t) count * sizeof(SdmPrimitive), cudaMemcpyHostToDevice));
    return 0;
}

// scene: 0 = the unmodified reference kernels (sd_obj); 1 = template<sd_obj>; 2 = template<sd_unit_mandelbulb>;
//        3 = template<SdTable> over the table of refgpu_set_table
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
    else if (scene == 2) tpl_refine<<<grid_for(n), BLOCK_SIZE>>>(in, out, SdUnitMandelbulb());
    else tpl_refine<<<grid_for(n), BLOCK_SIZE>>>(in, out, SdTable { g_table, g_table_count });
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
    else if (scene == 2) tpl_mesh<<<grid_for(n), BLOCK_SIZE>>>(in, d_tri, SdUnitMandelbulb());
    else tpl_mesh<<<grid_for(n), BLOCK_SIZE>>>(in, d_tri, SdTable { g_table, g_table_count });
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
    else if (scene == 3) tpl_sdf<<<grid_for(n), BLOCK_SIZE>>>(d_in, n, d_out, SdTable { g_table, g_table_count });
    else tpl_sdf<<<grid_for(n), BLOCK_SIZE>>>(d_in, n, d_out, SdObj());
    RCK(cudaGetLastError());
    RCK(cudaMemcpy(out, d_out, (size_t) n * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_out);
    return 0;
}

}  // extern "C"
