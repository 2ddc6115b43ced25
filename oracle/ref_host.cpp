// TEST INFRASTRUCTURE ONLY (oracle/_ref): the reference's own mesh-generation translation unit,
// compiled UNMODIFIED and BY PATH for the host CPU, behind a small C ABI.
//
//   REF_MESH_CU = "/root/reference/cuda/modules/compute_mesh_generation.cu"  (set by the Makefile)
//
// One loop iteration plays one CUDA thread (threadIdx/blockIdx are thread-locals supplied by
// glm_shim/host_shim.h); OpenMP runs "blocks" in parallel.  Built with -ffp-contract=off so that
// every float operation is a single IEEE-754 binary32 operation, which is what the product's
// kernels (nvcc -fmad=false, IEEE div/sqrt, no FTZ) reproduce bit for bit.
//
// Nothing here is shipped or measured as product; only tests/, smoke() and bench.py's CPU
// baseline leg may load the resulting oracle/_ref/libref_host.so.
#include REF_MESH_CU

#include <cstddef>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

// the two kernel bodies as templates over the SDF functor + the primitive-table functor (reference functions only)
#include "ref_functor.inc"

namespace {
template <class Body> void run_grid(unsigned n_threads_total, Body body) {
    // one loop iteration = one CUDA thread; OpenMP hands out small runs of consecutive threads, so that a launch of a few
    // hundred voxels still uses every host core (a block-granular split left the CPU arm's bounded samples single-threaded)
    const unsigned nb = (n_threads_total + BLOCK_SIZE - 1) / BLOCK_SIZE;
    const long long total = (long long) nb * BLOCK_SIZE;
#pragma omp parallel for schedule(dynamic, 8)
    for (long long g = 0; g < total; g++) {
        blockDim = { BLOCK_SIZE, 1, 1 };
        gridDim = { nb, 1, 1 };
        blockIdx = { (unsigned) (g / BLOCK_SIZE), 0, 0 };
        threadIdx = { (unsigned) (g % BLOCK_SIZE), 0, 0 };
        body();
    }
}
}  // namespace

extern "C" {

int ref_block_size() { return BLOCK_SIZE; }
int ref_init_factor() { return MESH_GENERATION_INIT_FACTOR; }
float ref_bb_size() { return MESH_GENERATION_BB_SIZE; }
int ref_sizeof_point() { return (int) sizeof(Point); }
int ref_sizeof_voxel_field() { return (int) sizeof(VoxelField); }
int ref_offsetof_voxels() { return (int) offsetof(VoxelField, voxels); }
int ref_offsetof_voxel_count() { return (int) offsetof(VoxelField, voxel_count); }
int ref_sizeof_vertex() { return (int) sizeof(Vertex); }
int ref_sizeof_triangle() { return (int) sizeof(Triangle); }
int ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// compute_mesh_generation.cu:12-62, launched as src/cuda/mod.rs:149-177 does (n threads, out count 8n)
void ref_refine(const float* voxels, unsigned n, const float* voxel_size, float* out_voxels /* 8n*3 */) {
    VoxelField in { { voxel_size[0], voxel_size[1], voxel_size[2] }, (Point*) voxels, n };
    VoxelField out { { 0.0f, 0.0f, 0.0f }, (Point*) out_voxels, n * 8u };
    run_grid(n, [&] { compute_refine_voxel_field_by_sdf(in, out); });
}

// compute_mesh_generation.cu:64-120, launched as src/cuda/mod.rs:226-250 does (5 Triangle slots / voxel)
void ref_mesh(const float* voxels, unsigned n, const float* voxel_size, float* out_triangles /* 5n*18 */) {
    VoxelField in { { voxel_size[0], voxel_size[1], voxel_size[2] }, (Point*) voxels, n };
    run_grid(n, [&] { compute_surface_triangles_from_voxel_field_by_sdf(in, (Triangle*) out_triangles); });
}

// ---- functor templates (ref_functor.inc): scene 1 = sd_obj (must equal the unmodified kernels above byte for byte),
//      scene 2 = sd_unit_mandelbulb, scene 3 = the SdmPrimitive table fold (e.g. the 1024-primitive scene), un-culled ----
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}
void ref_tpl_refine(int scene, const SdmPrimitive* prims, unsigned count, const float* voxels, unsigned n, const float* voxel_size,
                    float* out_voxels /* 8n*3 */) {
    VoxelField in { { voxel_size[0], voxel_size[1], voxel_size[2] }, (Point*) voxels, n };
    VoxelField out { { 0.0f, 0.0f, 0.0f }, (Point*) out_voxels, n * 8u };
    if (scene == 1) run_grid(n, [&] { tpl_refine(in, out, SdObj()); });
    else if (scene == 2) run_grid(n, [&] { tpl_refine(in, out, SdUnitMandelbulb()); });
    else run_grid(n, [&] { tpl_refine(in, out, SdTable { prims, count }); });
}
void ref_tpl_mesh(int scene, const SdmPrimitive* prims, unsigned count, const float* voxels, unsigned n, const float* voxel_size,
                  float* out_triangles /* 5n*18 */) {
    VoxelField in { { voxel_size[0], voxel_size[1], voxel_size[2] }, (Point*) voxels, n };
    if (scene == 1) run_grid(n, [&] { tpl_mesh(in, (Triangle*) out_triangles, SdObj()); });
    else if (scene == 2) run_grid(n, [&] { tpl_mesh(in, (Triangle*) out_triangles, SdUnitMandelbulb()); });
    else run_grid(n, [&] { tpl_mesh(in, (Triangle*) out_triangles, SdTable { prims, count }); });
}
void ref_tpl_sdf(const SdmPrimitive* prims, unsigned count, const float* p, unsigned n, float* out) {
    const SdTable sd { prims, count };
#pragma omp parallel for
    for (long long i = 0; i < (long long) n; i++) out[i] = sd(vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}

// ---- primitive-level probes (pin the oracle's restatement function by function) ------------
void ref_sd_obj(const float* p, unsigned n, float* out) {
#pragma omp parallel for
    for (long long i = 0; i < (long long) n; i++) out[i] = sd_obj(vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
float ref_smooth_min(float a, float b, float k) { return smooth_min(a, b, k); }
void ref_sd_box(const float* p, unsigned n, const float* bp, const float* bs, float* out) {
    for (unsigned i = 0; i < n; i++)
        out[i] = sd_box(vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), vec3(bp[0], bp[1], bp[2]), vec3(bs[0], bs[1], bs[2]));
}
void ref_sd_line(const float* p, unsigned n, const float* b0, const float* b1, float* out) {
    for (unsigned i = 0; i < n; i++)
        out[i] = sd_line(vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), vec3(b0[0], b0[1], b0[2]), vec3(b1[0], b1[1], b1[2]));
}
void ref_sd_box_skeleton(const float* p, unsigned n, const float* bp, const float* bs, float lw, float* out) {
    for (unsigned i = 0; i < n; i++)
        out[i] = sd_box_skeleton(vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), vec3(bp[0], bp[1], bp[2]),
                                 vec3(bs[0], bs[1], bs[2]), lw);
}
void ref_sd_unit_sphere(const float* p, unsigned n, float* out) {
    for (unsigned i = 0; i < n; i++) out[i] = sd_unit_sphere(vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
void ref_sd_unit_mandelbulb(const float* p, unsigned n, float* out) {
#pragma omp parallel for
    for (long long i = 0; i < (long long) n; i++) out[i] = sd_unit_mandelbulb(vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
void ref_empirical_normal_sd_obj(const float* p, unsigned n, float* out) {
#pragma omp parallel for
    for (long long i = 0; i < (long long) n; i++) {
        vec3 r = empirical_normal(sd_obj, vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
void ref_closest_surface_point_sd_obj(const float* p, unsigned n, float* out) {
#pragma omp parallel for
    for (long long i = 0; i < (long long) n; i++) {
        vec3 r = closest_surface_point(sd_obj, vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
// marching_cubes.cu:18-43 on one cube: values[8], vertices[8*3] -> triangles (<=5*18 floats, positions only
// are meaningful; normals are whatever Vertex{} default-initialises to), returns the count.
unsigned ref_march_cube(const float* values, const float* vertices, float* out_triangles) {
    McCube cube;
    for (int c = 0; c < 8; c++) {
        cube.values[c] = values[c];
        cube.vertices[c] = vec3(vertices[3 * c], vertices[3 * c + 1], vertices[3 * c + 2]);
    }
    Triangle tris[5];
    std::memset(tris, 0, sizeof(tris));
    unsigned n = march_cube(cube, tris);
    std::memcpy(out_triangles, tris, sizeof(Triangle) * n);
    return n;
}
void ref_mc_tables(int* edge_table /* 24 */, int* triangle_table /* 4096 */) {
    std::memcpy(edge_table, MC_EDGE_TABLE, sizeof(int) * 24);
    std::memcpy(triangle_table, MC_TRIANGLE_TABLE, sizeof(int) * 4096);
}

}  // extern "C"
