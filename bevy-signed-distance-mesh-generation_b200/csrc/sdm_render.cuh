// sdm_render.cuh - the ray-march viewer kernel: compute_render of the reference
// (cuda/modules/compute_render.cu:21-97) over the scene-table SDF of this library.
//
// One lane per pixel, pixels laid out as the reference lays them out (4 x 8 pixel tiles per warp, 2 x 2 warps per 128-thread
// block, cuda/modules/common.cu:186-215) so that a warp's rays stay coherent; the scene is the handle's primitive table -
// staged in shared memory for small tables, read through the per-cell primitive masks for large ones (the rays march through
// space, so every step looks its cell up by position; exact for the same reason as in the meshing kernels).  Arithmetic follows
// the reference statement by statement under the IEEE flags of this library (-fmad=false, IEEE division / sqrt, no FTZ); GLM
// forms as in oracle/glm_shim (mix, normalize, mat3 * vec3 by columns).  The reference's scene, sd_scene (compute_render.cu:3-19)
// = min(sd_obj, box skeleton of the meshing domain, lw 0.05), is the table { BOX_SKELETON, SPHERE smooth 0.5, BOX_SKELETON min }
// (bsdmg_b200.scenes.render_scene()): a min-fold of a skeleton's twelve edges into the accumulator is min(acc, min(edges)).
#pragma once

#include "sdm_device.cuh"

namespace sdm {

// bindings.h:16-29 (by-value kernel parameters of the reference; same layouts)
struct RenderGlobals { unsigned long long tick; float time; unsigned int render_texture_size[2]; float render_screen_size[2]; };
struct RenderCamera { float position[3]; float forward[3]; float up[3]; float right[3]; float fov; };
static_assert(sizeof(RenderGlobals) == 32 && sizeof(RenderCamera) == 52, "GlobalsBuffer / CameraBuffer layout (bindings.h:16-29)");

#define SDM_RAY_MARCH_STEP_LIMIT 256           /* ray_marching.cu:10 */
#define SDM_RAY_MARCH_DEPTH_LIMIT 500.0f       /* :11 */
#define SDM_RAY_MARCH_COLLISION_DISTANCE 0.001f /* :12 */
#define SDM_SQRT_INV 0.7071067811865475f       /* utils.cu:14 */

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
// glm::normalize: v * inversesqrt(dot(v, v)), inversesqrt = 1 / sqrt
__device__ __forceinline__ F3 normalize3(F3 v) {
    const float inv = 1.0f / sqrtf(dot3(v.x, v.y, v.z, v.x, v.y, v.z));
    return f3(v.x * inv, v.y * inv, v.z * inv);
}
__device__ __forceinline__ float length3(F3 v) { return sqrtf(dot3(v.x, v.y, v.z, v.x, v.y, v.z)); }

// common.cu:15-17, 69-74, 76-91: texture coordinate (pixel centre) -> ray direction.  `px, py` are the float texture coordinates.
__device__ __forceinline__ F3 pixel_to_dir(float px, float py, float tsx, float tsy, const RenderCamera& cam, float screen_x, float screen_y,
                                           float gtex_x, float gtex_y) {
    const float nx = (px + 0.5f) / tsx, ny = (py + 0.5f) / tsy;                       // texture_to_ndc
    const float cx = (2.0f * nx - 1.0f) * (tsx / tsy), cy = 1.0f - 2.0f * ny;         // ndc_to_camera (the texture size is passed as screen size)
    const float width_factor = (screen_x / gtex_x) * (gtex_y / screen_y);             // camera_to_ray
    const float fov_fac = tanf(cam.fov / 2.0f);
    const float a = cy * fov_fac, b = cx * fov_fac * width_factor;
    const F3 v = f3((cam.forward[0] + a * cam.up[0]) + b * cam.right[0], (cam.forward[1] + a * cam.up[1]) + b * cam.right[1],
                    (cam.forward[2] + a * cam.up[2]) + b * cam.right[2]);
    return normalize3(v);
}

// color.cu:7-21 (mat3x3 * vec3 = col0 * v.x + col1 * v.y + col2 * v.z, left to right)
__device__ __forceinline__ F3 aces_tone(F3 h) {
    const F3 v = f3((0.59719f * h.x + 0.35458f * h.y) + 0.04823f * h.z, (0.07600f * h.x + 0.90834f * h.y) + 0.01566f * h.z,
                    (0.02840f * h.x + 0.13383f * h.y) + 0.83777f * h.z);
    const F3 a = f3(v.x * (v.x + 0.0245786f) - 0.000090537f, v.y * (v.y + 0.0245786f) - 0.000090537f, v.z * (v.z + 0.0245786f) - 0.000090537f);
    const F3 b = f3(v.x * (0.983729f * v.x + 0.4329510f) + 0.238081f, v.y * (0.983729f * v.y + 0.4329510f) + 0.238081f,
                    v.z * (0.983729f * v.z + 0.4329510f) + 0.238081f);
    const F3 q = f3(a.x / b.x, a.y / b.y, a.z / b.z);
    const F3 r = f3((1.60475f * q.x + -0.53108f * q.y) + -0.07367f * q.z, (-0.10208f * q.x + 1.10813f * q.y) + -0.00605f * q.z,
                    (-0.00327f * q.x + -0.07276f * q.y) + 1.07602f * q.z);
    // min(max(r, vec3(0)), vec3(1)) with glm's component forms: max(x, y) = (x < y) ? y : x, min(x, y) = (y < x) ? y : x
    auto sat = [](float x) { const float m = (x < 0.0f) ? 0.0f : x; return (1.0f < m) ? 1.0f : m; };
    return f3(sat(r.x), sat(r.y), sat(r.z));
}

// one pixel per lane; shared by the library kernel (k_render) and the reference-ABI module (csrc/compat_render.cu)
__device__ __forceinline__ void render_pixel(const SceneView& sc, const MaskGrid& grid, uchar4* __restrict__ out, const RenderGlobals& g, const RenderCamera& cam,
                                             uint32_t tex_w, uint32_t tex_h) {
    // render_texture_coord (common.cu:186-215): 4 x 8 pixels per warp, 2 x 2 warps per block, blocks row-major over width / 8 columns
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bcx = (int) tex_w / 8;
    const int tx = 8 * ((int) blockIdx.x % bcx) + (lane % 4) + 4 * (warp % 2);
    const int ty = 16 * ((int) blockIdx.x / bcx) + (lane / 4) + 8 * (warp / 2);
    const bool inside = (uint32_t) ty < tex_h && (uint32_t) tx < tex_w;   // compute_render.cu:30-32 (the reference returns here)
    const float tsx = (float) tex_w, tsy = (float) tex_h;
    const float gtx = (float) g.render_texture_size[0], gty = (float) g.render_texture_size[1];
    const float sx = g.render_screen_size[0], sy = g.render_screen_size[1];
    const F3 dir = pixel_to_dir((float) tx, (float) ty, tsx, tsy, cam, sx, sy, gtx, gty);
    // get_pixel_cone_radius (common.cu:95-184): widest gap between the ray and the rays through the pixel's four "corners"
    float cone = 0.0f;
    {
        const float ox[4] = { -SDM_SQRT_INV, -SDM_SQRT_INV, SDM_SQRT_INV, SDM_SQRT_INV }, oy[4] = { -SDM_SQRT_INV, SDM_SQRT_INV, -SDM_SQRT_INV, SDM_SQRT_INV };
        float l[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const F3 b = pixel_to_dir((float) tx + ox[q], (float) ty + oy[q], tsx, tsy, cam, sx, sy, gtx, gty);
            l[q] = length3(f3(dir.x - b.x, dir.y - b.y, dir.z - b.z));
        }
        cone = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
    }
    // ray_march (ray_marching.cu:14-49)
    float px = cam.position[0], py = cam.position[1], pz = cam.position[2], depth = 0.0f;
    int outcome = 1;   // StepLimit
    bool running = inside;
    for (int step = 0; step < SDM_RAY_MARCH_STEP_LIMIT; step++) {
        if (!__any_sync(0xffffffffu, running)) break;
        tile_mask_from_point(grid, sc, running, px, py, pz);   // culled scenes: the lanes' cells (the evaluation point itself, no stencil needed)
        if (running) {
            const float collision_distance = cone * depth;
            const float d = eval_scene1(sc, px, py, pz);
            if (d <= collision_distance + SDM_RAY_MARCH_COLLISION_DISTANCE) { outcome = 0; running = false; }
            else {
                const float adv = d - collision_distance;
                depth += adv;
                px += adv * dir.x; py += adv * dir.y; pz += adv * dir.z;
                if (depth > SDM_RAY_MARCH_DEPTH_LIMIT) { outcome = 2; running = false; }
            }
        }
    }
    // shading (compute_render.cu:64-88)
    const bool hit = inside && outcome == 0;
    F3 color = f3(0.0f, 0.0f, 0.0f);
    if (__any_sync(0xffffffffu, hit)) {
        tile_mask_from_point(grid, sc, hit, px, py, pz);
        if (hit) {
            float nx, ny, nz;
            empirical_normal(sc, px, py, pz, nx, ny, nz);
            const F3 light = normalize3(f3(1.0f, 1.0f, 1.0f));
            const float a = (dot3(nx, ny, nz, light.x, light.y, light.z) + 1.0f) / 2.0f;
            const F3 c0 = f3(19.0f / 255.0f, 9.0f / 255.0f, 130.0f / 255.0f), c1 = f3(240.0f / 255.0f, 103.0f / 255.0f, 24.0f / 255.0f);
            color = f3(c0.x * (1.0f - a) + c1.x * a, c0.y * (1.0f - a) + c1.y * a, c0.z * (1.0f - a) + c1.z * a);   // glm::mix
        }
    }
    if (inside && outcome == 1) color = f3(1.0f, 1.0f, 1.0f);
    if (!inside) return;
    color = aces_tone(color);
    auto clamp01 = [](float x) { const float m = (x < 0.0f) ? 0.0f : x; return (1.0f < m) ? 1.0f : m; };   // glm::clamp = min(max(x, lo), hi)
    // index_2d (common.cu:32-35)
    const uint32_t idx = (uint32_t) min(max(tx, 0), (int) tex_w - 1) + (uint32_t) min(max(ty, 0), (int) tex_h - 1) * tex_w;
    out[idx] = make_uchar4((unsigned char) (clamp01(color.x) * 255.0f), (unsigned char) (clamp01(color.y) * 255.0f), (unsigned char) (clamp01(color.z) * 255.0f), 0xFF);
}

__global__ void __launch_bounds__(128) k_render(const uint4* __restrict__ scene, uchar4* __restrict__ out, RenderGlobals g, RenderCamera cam, uint32_t tex_w,
                                               uint32_t tex_h, MaskGrid grid) {
    extern __shared__ uint4 smem[];
    const SceneView sc = stage_scene_masked(scene, smem, grid);
    render_pixel(sc, grid, out, g, cam, tex_w, tex_h);
}

}  // namespace sdm
