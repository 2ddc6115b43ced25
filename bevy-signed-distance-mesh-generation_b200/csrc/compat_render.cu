// compat_render.cu - the reference's ray-march module at kernel level, for sm_100a.
//
// The reference's Rust host loads the PTX module "compute_render" and launches `compute_render` by name with three by-value
// #[repr(C)] structs (src/cuda/mod.rs:70-79, 86-90, 372-399; layouts cuda/includes/bindings.h:16-41): grid = w * h / 128 blocks of
// 128 threads, no dynamic shared memory.  This translation unit exports that symbol with those parameter layouts, so the UNMODIFIED
// host's viewer runs on a B200 by swapping assets/cuda/compiled/compute_render.ptx for the file built from this source
// (csrc/Makefile: `make compat`).  The scene is the reference's compiled-in sd_scene (cuda/modules/compute_render.cu:3-19):
// min(sd_obj, box skeleton of the meshing domain with lw 0.05), built in shared memory as a 25-primitive table; the pixel code is
// the library's (csrc/sdm_render.cuh), bit-identical to the reference kernel compiled with IEEE arithmetic
// (tests/test_gpu_compat.py loads the module through the CUDA driver API, as cudarc does).
#include "sdm_render.cuh"

using namespace sdm;

extern "C" {
struct Rgba { unsigned char r, g, b, a; };
struct RenderTexture { unsigned int size[2]; Rgba* data; };
struct GlobalsBuffer { unsigned long long tick; float time; unsigned int render_texture_size[2]; float render_screen_size[2]; };
struct CameraBuffer { float position[3]; float forward[3]; float up[3]; float right[3]; float fov; };
}
static_assert(sizeof(RenderTexture) == 16 && sizeof(GlobalsBuffer) == 32 && sizeof(CameraBuffer) == 52, "bindings.h layouts");

namespace {

struct SdSceneTable { SceneHeader hdr; DevRun runs[3]; DevPrim prims[25]; };

__device__ __forceinline__ void put_capsule(DevPrim& d, float ax, float ay, float az, float bx, float by, float bz, float lw) {
    const float ex = bx - ax, ey = by - ay, ez = bz - az;
    const float len = sqrtf(ex * ex + ey * ey + ez * ez);          // length(b1 - b0)          (signed_distance.cu:78)
    d.v0[0] = ax; d.v0[1] = ay; d.v0[2] = az;
    d.v1[0] = ex / len; d.v1[1] = ey / len; d.v1[2] = ez / len;    // (b1 - b0) / len          (:79)
    d.v2[0] = 0.f; d.v2[1] = 0.f; d.v2[2] = 0.f;
    d.s0 = lw; d.s1 = len; d.k = 0.0f;
    d.kind = SDM_PRIM_CAPSULE; d.fold = SDM_FOLD_MIN; d.pad0 = 0; d.pad1 = 0;
}
// the twelve edges of sd_box_skeleton(p, 0, bs, lw) in the loop order of signed_distance.cu:97-99 (with its `% 2`, :101)
__device__ __forceinline__ void put_skeleton_edge(DevPrim& d, int t, float bsx, float bsy, float bsz, float lw) {
    const float bs[3] = { bsx, bsy, bsz };
    const float bpl[3] = { 0.0f - bs[0] / 2.0f, 0.0f - bs[1] / 2.0f, 0.0f - bs[2] / 2.0f };
    const int dir = t >> 2, c0 = (t >> 1) & 1, c1 = t & 1;
    float m0[3] = { bpl[0], bpl[1], bpl[2] };
    m0[(dir + 1) % 3] += c0 ? bs[(dir + 1) % 2] : 0.0f;
    m0[(dir + 2) % 3] += c1 ? bs[(dir + 2) % 3] : 0.0f;
    float m1[3] = { m0[0], m0[1], m0[2] };
    m1[dir] += bs[dir];
    put_capsule(d, m0[0], m0[1], m0[2], m1[0], m1[1], m1[2], lw);
}
// sd_scene = min(smooth_min(sd_box_skeleton(p, 0, (3,1,.5), .1), length(p) - 1, .5), sd_box_skeleton(p, 0, (5,5,5), .05))
__device__ __forceinline__ SceneView build_sd_scene(SdSceneTable* s) {
    const int t = threadIdx.x;
    if (t < 12) put_skeleton_edge(s->prims[t], t, 3.0f, 1.0f, 0.5f, 0.1f);
    else if (t == 12) {
        DevPrim& d = s->prims[12];
        d.v0[0] = d.v0[1] = d.v0[2] = 0.0f; d.s0 = 1.0f;
        d.v1[0] = d.v1[1] = d.v1[2] = 0.0f; d.s1 = 0.0f;
        d.v2[0] = d.v2[1] = d.v2[2] = 0.0f; d.k = 0.5f;
        d.kind = SDM_PRIM_SPHERE; d.fold = SDM_FOLD_SMOOTH_MIN; d.pad0 = 0; d.pad1 = 0;
    } else if (t < 25) {
        // (MESH_GENERATION_BB_MIN + MESH_GENERATION_BB_MAX) / 2 = 0, MAX - MIN = 5 on every axis (common.cu:219-220)
        const float lo = 0.0f - (5.0f / 2.0f), hi = 5.0f / 2.0f;
        put_skeleton_edge(s->prims[t], t - 13, hi - lo, hi - lo, hi - lo, 0.05f);
    } else if (t == 25) {
        s->runs[0] = DevRun { SDM_PRIM_CAPSULE, SDM_FOLD_MIN, 0u, 12u | ((uint32_t) RUN_SHARED_RADIUS_MIN << 24) };
        s->runs[1] = DevRun { SDM_PRIM_SPHERE, SDM_FOLD_SMOOTH_MIN, 12u, 1u };
        s->runs[2] = DevRun { SDM_PRIM_CAPSULE, SDM_FOLD_MIN, 13u, 12u | ((uint32_t) RUN_SHARED_RADIUS_MIN << 24) };
    }
    __syncthreads();
    SceneView v;
    v.prims = s->prims; v.runs = s->runs; v.nruns = 3; v.nprims = 25;
    v.wmask = nullptr; v.W = 0; v.tlist = nullptr; v.tcount = nullptr;
    v.kmax = 0.5f;
    return v;
}

}  // namespace

extern "C" __global__ void __launch_bounds__(128) compute_render(const RenderTexture render_texture, const GlobalsBuffer globals, const CameraBuffer camera) {
    __shared__ SdSceneTable scene;
    const SceneView sc = build_sd_scene(&scene);
    RenderGlobals g;
    g.tick = globals.tick; g.time = globals.time;
    g.render_texture_size[0] = globals.render_texture_size[0]; g.render_texture_size[1] = globals.render_texture_size[1];
    g.render_screen_size[0] = globals.render_screen_size[0]; g.render_screen_size[1] = globals.render_screen_size[1];
    RenderCamera c;
    for (int i = 0; i < 3; i++) { c.position[i] = camera.position[i]; c.forward[i] = camera.forward[i]; c.up[i] = camera.up[i]; c.right[i] = camera.right[i]; }
    c.fov = camera.fov;
    MaskGrid grid;
    grid.masks = nullptr; grid.G = 0; grid.W = 0; grid.ox = grid.oy = grid.oz = 0.0f; grid.cell = 0.0f; grid.inv_cell = 0.0f; grid.enabled = 0; grid.maybe = nullptr;
    render_pixel(sc, grid, reinterpret_cast<uchar4*>(render_texture.data), g, c, render_texture.size[0], render_texture.size[1]);
}
