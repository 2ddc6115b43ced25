// TEST INFRASTRUCTURE ONLY (oracle/_ref): the reference's ray-march module, cuda/modules/compute_render.cu, compiled UNMODIFIED and
// BY PATH for sm_100a with the IEEE flags of this repo (-fmad=false, IEEE division / sqrt, no FTZ; the reference's own build uses
// --use_fast_math, build.rs:113-118), behind a launcher with a C ABI that launches it exactly as src/cuda/mod.rs:382-399 does
// (grid = w * h / BLOCK_SIZE, block = BLOCK_SIZE, by-value RenderTexture / GlobalsBuffer / CameraBuffer).
#include REF_RENDER_CU

#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#define RCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "ref_render: %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

extern "C" {
int refrender_sizeof_globals() { return (int) sizeof(GlobalsBuffer); }
int refrender_sizeof_camera() { return (int) sizeof(CameraBuffer); }
int refrender(const void* globals, const void* camera, unsigned char* out_rgba) {
    GlobalsBuffer g;
    CameraBuffer c;
    memcpy(&g, globals, sizeof g);
    memcpy(&c, camera, sizeof c);
    const unsigned w = g.render_texture_size[0], h = g.render_texture_size[1];
    Rgba* d = nullptr;
    RCK(cudaMalloc(&d, (size_t) w * h * 4));
    RCK(cudaMemset(d, 0, (size_t) w * h * 4));
    RenderTexture t;
    t.size[0] = w; t.size[1] = h; t.data = d;
    compute_render<<<(w * h) / BLOCK_SIZE, BLOCK_SIZE>>>(t, g, c);
    RCK(cudaGetLastError());
    RCK(cudaMemcpy(out_rgba, d, (size_t) w * h * 4, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}
}
