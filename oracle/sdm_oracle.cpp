// ============================================================================================
// TEST INFRASTRUCTURE ONLY - CPU oracle for the mesh-generation hot path.
//
// A plain C++ restatement (scalar float code, no GLM, no CUDA) of the reference's algorithm:
//   cuda/includes/signed_distance.cu      SDF primitives, smooth_min, empirical_normal, closest_surface_point
//   cuda/includes/marching_cubes.cu       march_cube (fixed mid-point edge vertices)
//   cuda/modules/common.cu:222-226        sd_obj
//   cuda/modules/compute_mesh_generation.cu:12-62, 64-120   the two kernels
//   src/cuda/mod.rs:105-122, 179-194, 263-296                level-0 field, stable retain, vertex weld
// Every function cites the lines it follows.  The scene is the SdmPrimitive fold of include/sdfmesh.h
// so that the same scene table can be handed to the oracle and to the CUDA library.
//
// Parity pinning: the reference has no tests or golden vectors.  This file is
// pinned instead against the reference's OWN code compiled unmodified for the host
// (oracle/_ref/libref_host.so, built by oracle/Makefile from /root/reference) - tests/test_oracle_vs_ref.py
// requires byte equality on sd_obj at levels 0-3 - and against the golden hashes under tests/golden/
// that were generated with that library (tools/gen_golden.py).  The GLM arithmetic the reference
// relies on is restated from GLM 1.0.1's published formulas (glm_shim/...) because GLM itself is not
// vendored: that part is "parity unpinned" against real GLM.
//
// Build: g++ -O2 -ffp-contract=off (one IEEE-754 binary32 operation per source operation; no FMA).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  The product (libsdfmesh.so) never does.
// ============================================================================================
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/sdfmesh.h"
#include "mc_tables_oracle.inc"

namespace {

struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { return V3 { x, y, z }; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline float& at(V3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
inline float at(const V3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// GLM 1.0.1 semantics (see glm_shim/includes/libraries/glm/glm.hpp for the formulas' provenance)
inline float g_dot(V3 a, V3 b) { V3 t = a * b; return t.x + t.y + t.z; }
inline float g_length(V3 v) { return sqrtf(g_dot(v, v)); }
inline float g_distance(V3 p0, V3 p1) { return g_length(p1 - p0); }
inline V3 g_normalize(V3 v) { return v * (1.0f / sqrtf(g_dot(v, v))); }
inline V3 g_cross(V3 x, V3 y) { return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline float g_absf(float x) { return x >= 0.0f ? x : -x; }          // glm::abs on vec components
inline float g_minf(float x, float y) { return (y < x) ? y : x; }    // glm::min on vec components
inline float g_maxf(float x, float y) { return (x < y) ? y : x; }    // glm::max on vec components

// signed_distance.cu:20-23 (scalar abs/min/max resolve to CUDA's fabsf/fminf/fmaxf)
inline float smooth_min(float a, float b, float k) {
    float h = fmaxf(k - fabsf(a - b), 0.0f) / k;
    return fminf(a, b) - h * h * h * k * (1.0f / 6.0f);
}

// signed_distance.cu:65-75
inline float sd_ray(V3 p, V3 bl, V3 bd, float len) {
    float d = g_dot(p - bl, bd);
    if (d < 0) {
        return g_distance(bl, p);
    } else if (d > len) {
        return g_distance(bl + len * bd, p);
    }
    return g_distance(bl + bd * d, p);
}
// signed_distance.cu:77-80
inline float sd_line(V3 p, V3 b0, V3 b1) {
    float len = g_length(b1 - b0);
    return sd_ray(p, b0, (b1 - b0) / len, len);
}
// signed_distance.cu:86-91 (utils.cu:20 maximum)
inline float sd_box(V3 p, V3 bp, V3 bs) {
    V3 d = p - bp;
    V3 q = v3(g_absf(d.x), g_absf(d.y), g_absf(d.z)) - bs / 2.0f;
    float udst = g_length(v3(g_maxf(q.x, 0.0f), g_maxf(q.y, 0.0f), g_maxf(q.z, 0.0f)));
    V3 m = v3(g_minf(q.x, 0.0f), g_minf(q.y, 0.0f), g_minf(q.z, 0.0f));
    float idst = fmaxf(fmaxf(m.x, m.y), m.z);
    return udst + idst;
}
// signed_distance.cu:93-113, including the bs[(dir + 1) % 2] indexing of :101
inline float sd_box_skeleton(V3 p, V3 bp, V3 bs, float lw) {
    V3 bpl = bp - bs / 2.0f;
    float sd = (float) 3.40282347E+38;  // utils.cu:10 MAX_POSITIVE_F32
    for (int dir = 0; dir < 3; dir++) {
        for (int c0 = 0; c0 < 2; c0++) {
            for (int c1 = 0; c1 < 2; c1++) {
                V3 m0 = bpl;
                at(m0, (dir + 1) % 3) += c0 ? at(bs, (dir + 1) % 2) : 0.0f;
                at(m0, (dir + 2) % 3) += c1 ? at(bs, (dir + 2) % 3) : 0.0f;
                V3 m1 = m0;
                at(m1, dir) += at(bs, dir);
                sd = fminf(sd, sd_line(p, m0, m1) - lw);
            }
        }
    }
    return sd;
}
// signed_distance.cu:27-53 with time = 0 (and :55-57 for the scaled wrapper).  float overloads of the
// libm functions (CUDA device code resolves acos/atan2/pow/sin/cos/log on float to the *f forms).
// CPU libm and CUDA libdevice differ in the last ulp, so this scene is NOT bit-comparable CPU<->GPU;
// tests compare the GPU product against the reference kernels compiled for the GPU instead.
inline float sd_mandelbulb0(V3 p) {
    V3 z = p;
    float dr = 1.0f;
    float r = 0.0f;
    float power = 7.0f * (1.0f + 0.0f * 0.001f);
    for (int i = 0; i < 25; i++) {
        r = g_length(z);
        if (r > 2.0f) break;
        float theta = acosf(z.z / r) * power;
        float phi = atan2f(z.y, z.x) * power;
        float zr = powf(r, power);
        dr = powf(r, power - 1.0f) * power * dr + 1.0f;
        float s_theta = sinf(theta);
        z = zr * v3(s_theta * cosf(phi), sinf(phi) * s_theta, cosf(theta));
        z = z + p;
    }
    return 0.5f * logf(r) * r / dr;
}

struct Scene {
    const SdmPrimitive* prims;
    uint32_t count;
    // include/sdfmesh.h scene fold: acc = FLT_MAX; acc = fold_i(acc, d_i(p)) in index order.
    float operator()(V3 p) const {
        float acc = (float) 3.40282347E+38;
        for (uint32_t i = 0; i < count; i++) {
            const SdmPrimitive& q = prims[i];
            V3 a = v3(q.a[0], q.a[1], q.a[2]), b = v3(q.b[0], q.b[1], q.b[2]);
            float d;
            switch (q.kind) {
                case SDM_PRIM_SPHERE: d = g_length(p - a) - q.radius; break;              // common.cu:224
                case SDM_PRIM_BOX: d = sd_box(p, a, b); break;
                case SDM_PRIM_CAPSULE: d = sd_line(p, a, b) - q.radius; break;            // signed_distance.cu:109
                case SDM_PRIM_BOX_SKELETON: d = sd_box_skeleton(p, a, b, q.radius); break;
                case SDM_PRIM_MANDELBULB: d = sd_mandelbulb0(p / q.radius) * q.radius; break;  // :55-57
                default: d = NAN;
            }
            acc = (q.fold == SDM_FOLD_SMOOTH_MIN) ? smooth_min(acc, d, q.k) : fminf(acc, d);
        }
        return acc;
    }
};

// signed_distance.cu:179-202
const float NORMAL_EPSILON = 0.001f;
template <class F> inline V3 empirical_normal(const F& sd, V3 p) {
    float dx = (-sd(p + v3(2.0f * NORMAL_EPSILON, 0.0f, 0.0f)) + 8.0f * sd(p + v3(NORMAL_EPSILON, 0.0f, 0.0f)) -
                8.0f * sd(p + v3(-NORMAL_EPSILON, 0.0f, 0.0f)) + sd(p + v3(-2.0f * NORMAL_EPSILON, 0.0f, 0.0f)));
    float dy = (-sd(p + v3(0.0f, 2.0f * NORMAL_EPSILON, 0.0f)) + 8.0f * sd(p + v3(0.0f, NORMAL_EPSILON, 0.0f)) -
                8.0f * sd(p + v3(0.0f, -NORMAL_EPSILON, 0.0f)) + sd(p + v3(0.0f, -2.0f * NORMAL_EPSILON, 0.0f)));
    float dz = (-sd(p + v3(0.0f, 0.0f, 2.0f * NORMAL_EPSILON)) + 8.0f * sd(p + v3(0.0f, 0.0f, NORMAL_EPSILON)) -
                8.0f * sd(p + v3(0.0f, 0.0f, -NORMAL_EPSILON)) + sd(p + v3(0.0f, 0.0f, -2.0f * NORMAL_EPSILON)));
    return g_normalize(v3(dx, dy, dz));
}
// signed_distance.cu:227-240
template <class F> inline V3 closest_surface_point(const F& sd_func, V3 p, uint32_t* iters = nullptr) {
    V3 g = p;
    bool collision = false;
    int i = 0;
    for (; !collision && i < 10000; i++) {
        float sd = sd_func(g);
        V3 n = empirical_normal(sd_func, g);
        g = g - sd * n;
        collision = fabsf(sd) <= 0.00001f;
    }
    if (iters) *iters = (uint32_t) i;
    return g;
}

inline V3 ld(const float* p, size_t i) { return v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline void st(float* p, size_t i, V3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

}  // namespace

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}

void orc_eval_sdf(const SdmPrimitive* prims, uint32_t count, const float* pts, uint32_t n, float* out) {
    Scene sc { prims, count };
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long) n; i++) out[i] = sc(ld(pts, i));
}
void orc_eval_normal(const SdmPrimitive* prims, uint32_t count, const float* pts, uint32_t n, float* out) {
    Scene sc { prims, count };
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long) n; i++) st(out, i, empirical_normal(sc, ld(pts, i)));
}
void orc_eval_project(const SdmPrimitive* prims, uint32_t count, const float* pts, uint32_t n, float* out,
                      uint32_t* out_iters) {
    Scene sc { prims, count };
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long) n; i++) {
        uint32_t it = 0;
        st(out, i, closest_surface_point(sc, ld(pts, i), &it));
        if (out_iters) out_iters[i] = it;
    }
}
float orc_smooth_min(float a, float b, float k) { return smooth_min(a, b, k); }

// src/cuda/mod.rs:105-122: x outer, y, z inner; SIZE = bb / init in f32; coordinate = i*SIZE - bb/2.
void orc_create_voxel_field(float bb_size, uint32_t init_factor, float* out_voxels, float* out_voxel_size) {
    const float size = bb_size / (float) init_factor;
    out_voxel_size[0] = out_voxel_size[1] = out_voxel_size[2] = size;
    size_t o = 0;
    for (uint32_t x = 0; x < init_factor; x++)
        for (uint32_t y = 0; y < init_factor; y++)
            for (uint32_t z = 0; z < init_factor; z++, o++) {
                out_voxels[3 * o + 0] = (float) x * size - bb_size / 2.0f;
                out_voxels[3 * o + 1] = (float) y * size - bb_size / 2.0f;
                out_voxels[3 * o + 2] = (float) z * size - bb_size / 2.0f;
            }
}

// compute_mesh_generation.cu:12-62.  out_voxels has 8n slots: child min-corner or (INF,INF,INF).
void orc_refine_raw(const SdmPrimitive* prims, uint32_t count, const float* voxels, uint32_t n, const float* voxel_size,
                    float* out_voxels) {
    Scene sc { prims, count };
    const V3 osz = v3(voxel_size[0], voxel_size[1], voxel_size[2]) / 2.0f;  // :20
#pragma omp parallel for schedule(dynamic, 64)
    for (long long id = 0; id < (long long) n; id++) {
        const V3 base = ld(voxels, id);
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 2; k++) {
                    V3 lower = base + v3((float) i, (float) j, (float) k) * osz;
                    V3 upper = base + v3((float) (i + 1), (float) (j + 1), (float) (k + 1)) * osz;
                    bool is_border = false;
                    bool prev = sc(lower) <= 0.0f;
                    for (int c = 1; c < 8; c++) {
                        V3 q = v3(c & 1 ? upper.x : lower.x, c & 2 ? upper.y : lower.y, c & 4 ? upper.z : lower.z);
                        if (prev != (sc(q) <= 0.0f)) { is_border = true; break; }
                    }
                    const size_t n_id = (size_t) id * 8 + i * 4 + j * 2 + k;
                    out_voxels[3 * n_id + 0] = is_border ? lower.x : INFINITY;
                    out_voxels[3 * n_id + 1] = is_border ? lower.y : INFINITY;
                    out_voxels[3 * n_id + 2] = is_border ? lower.z : INFINITY;
                }
    }
}
// src/cuda/mod.rs:192-193: Vec::retain(all three finite) - stable.  Returns the surviving count.
uint32_t orc_retain_finite(float* voxels, uint32_t n) {
    uint32_t w = 0;
    for (uint32_t i = 0; i < n; i++) {
        float x = voxels[3 * i], y = voxels[3 * i + 1], z = voxels[3 * i + 2];
        if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z)) {
            voxels[3 * w] = x; voxels[3 * w + 1] = y; voxels[3 * w + 2] = z;
            w++;
        }
    }
    return w;
}

// compute_mesh_generation.cu:64-120 (+ marching_cubes.cu:13-43).  out_triangles: 5n Triangle slots of 18
// floats, unused slots = {NaN x6, 0 x12}.  out_cases (optional): cube_index per voxel.
void orc_mesh_raw(const SdmPrimitive* prims, uint32_t count, const float* voxels, uint32_t n, const float* voxel_size,
                  float* out_triangles, uint8_t* out_cases) {
    Scene sc { prims, count };
    const V3 vs = v3(voxel_size[0], voxel_size[1], voxel_size[2]);
#pragma omp parallel for schedule(dynamic, 16)
    for (long long id = 0; id < (long long) n; id++) {
        const V3 base = ld(voxels, id);
        V3 cv[8];
        float val[8];
        for (int c = 0; c < 8; c++) {  // :77-86
            V3 v = base;
            v.x += ((c % 4) == 1 || (c % 4) == 2) ? vs.x : 0.0f;
            v.y += (c % 4) >= 2 ? vs.y : 0.0f;
            v.z += c >= 4 ? vs.z : 0.0f;
            cv[c] = v;
            val[c] = sc(v);
        }
        unsigned cube_index = 0;  // marching_cubes.cu:19-23
        for (int i = 0; i < 8; i++) cube_index |= ((unsigned) (val[i] <= 0.0f)) << i;
        if (out_cases) out_cases[id] = (uint8_t) cube_index;
        const char* row = ORC_MC_TRIANGLES[cube_index];
        const unsigned ntri = (unsigned) (strlen(row) / 3);
        float* out = out_triangles + (size_t) id * 5 * 18;
        for (unsigned t = 0; t < ntri; t++) {
            V3 v[3];
            for (int j = 0; j < 3; j++) {  // marching_cubes.cu:13-16: mix(a, b, 0.5f) = a*(1-0.5) + b*0.5
                char ch = row[3 * t + j];
                int e = (ch <= '9') ? ch - '0' : ch - 'a' + 10;
                v[j] = cv[ORC_MC_EDGES[e][0]] * (1.0f - 0.5f) + cv[ORC_MC_EDGES[e][1]] * 0.5f;
            }
            V3 v0 = closest_surface_point(sc, v[0]);  // :95-97
            V3 v1 = closest_surface_point(sc, v[1]);
            V3 v2 = closest_surface_point(sc, v[2]);
            V3 n0 = empirical_normal(sc, v0);  // :99-101
            V3 n1 = empirical_normal(sc, v1);
            V3 n2 = empirical_normal(sc, v2);
            const V3 triangle_normal = g_normalize(g_cross(v1 - v0, v2 - v0));          // :103
            const V3 actual_normal = empirical_normal(sc, (v0 + v1 + v2) / 3.0f);       // :104
            const bool flip = g_dot(triangle_normal, actual_normal) <= 0.0f;            // :105
            float* o = out + t * 18;
            st(o, 0, flip ? v2 : v0); st(o, 1, flip ? n2 : n0);                         // :107-113
            st(o, 2, v1);             st(o, 3, n1);
            st(o, 4, flip ? v0 : v2); st(o, 5, flip ? n0 : n2);
        }
        // :116-118 `triangles[i] = { POINT_NAN, POINT_NAN }`: brace elision makes that vertices[0] =
        // {position NaN, normal NaN} and value-initialises vertices[1..2] to zero (not NaN).
        for (unsigned t = ntri; t < 5; t++)
            for (int q = 0; q < 18; q++) out[t * 18 + q] = q < 6 ? NAN : 0.0f;
    }
}

// src/cuda/mod.rs:263-296.  Key = [(x * 10e4f32).round() as i64; 3]; Rust `round` is half-away-from-zero,
// `as i64` saturates and maps NaN to 0.  A triangle is kept iff vertices[0].position.x is finite.
// Outputs are sized for the worst case by the caller (3 * n_slots vertices); returns counts via pointers.
static inline long long rust_f32_to_i64(float v) {
    if (std::isnan(v)) return 0;
    if (v >= 9223372036854775808.0f) return INT64_MAX;
    if (v <= -9223372036854775808.0f) return INT64_MIN;
    return (long long) v;
}
struct Key3 { long long a, b, c; bool operator==(const Key3& o) const { return a == o.a && b == o.b && c == o.c; } };
struct Key3Hash {
    size_t operator()(const Key3& k) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (uint64_t v : { (uint64_t) k.a, (uint64_t) k.b, (uint64_t) k.c }) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
        return (size_t) h;
    }
};
void orc_weld(const float* triangles, uint32_t n_slots, float* out_positions, float* out_normals, uint32_t* out_indices,
              uint32_t* out_vertex_count, uint32_t* out_triangle_count) {
    std::unordered_map<Key3, uint32_t, Key3Hash> map;
    map.reserve((size_t) n_slots);
    uint32_t nv = 0, nt = 0;
    for (uint32_t s = 0; s < n_slots; s++) {
        const float* t = triangles + (size_t) s * 18;
        if (!std::isfinite(t[0])) continue;
        for (int j = 0; j < 3; j++) {
            const float* vtx = t + j * 6;
            Key3 key { rust_f32_to_i64(roundf(vtx[0] * 10e4f)), rust_f32_to_i64(roundf(vtx[1] * 10e4f)),
                       rust_f32_to_i64(roundf(vtx[2] * 10e4f)) };
            auto it = map.find(key);
            uint32_t idx;
            if (it == map.end()) {
                idx = nv++;
                map.emplace(key, idx);
                memcpy(out_positions + 3 * (size_t) idx, vtx, 12);
                memcpy(out_normals + 3 * (size_t) idx, vtx + 3, 12);
            } else {
                idx = it->second;
            }
            out_indices[3 * (size_t) nt + j] = idx;
        }
        nt++;
    }
    *out_vertex_count = nv;
    *out_triangle_count = nt;
}

// Newton iteration statistics of closest_surface_point over the raw edge mid-points (design input).
void orc_march_cube(const float* values, const float* vertices, float* out_positions /* 5*3*3 */, uint32_t* out_count,
                    uint8_t* out_case) {
    unsigned cube_index = 0;
    for (int i = 0; i < 8; i++) cube_index |= ((unsigned) (values[i] <= 0.0f)) << i;
    *out_case = (uint8_t) cube_index;
    const char* row = ORC_MC_TRIANGLES[cube_index];
    const unsigned ntri = (unsigned) (strlen(row) / 3);
    for (unsigned t = 0; t < ntri; t++)
        for (int j = 0; j < 3; j++) {
            char ch = row[3 * t + j];
            int e = (ch <= '9') ? ch - '0' : ch - 'a' + 10;
            V3 p = ld(vertices, ORC_MC_EDGES[e][0]) * (1.0f - 0.5f) + ld(vertices, ORC_MC_EDGES[e][1]) * 0.5f;
            st(out_positions, t * 3 + j, p);
        }
    *out_count = ntri;
}
uint64_t orc_fnv1a64(const void* data, size_t n) {
    const unsigned char* b = (const unsigned char*) data;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}
void orc_mc_tables(int* edge_table /* 24 */, int* triangle_table /* 4096 */) {
    for (int e = 0; e < 12; e++) { edge_table[2 * e] = ORC_MC_EDGES[e][0]; edge_table[2 * e + 1] = ORC_MC_EDGES[e][1]; }
    for (int c = 0; c < 256; c++) {
        const char* row = ORC_MC_TRIANGLES[c];
        int len = (int) strlen(row);
        for (int j = 0; j < 16; j++) {
            if (j < len) { char ch = row[j]; triangle_table[16 * c + j] = (ch <= '9') ? ch - '0' : ch - 'a' + 10; }
            else triangle_table[16 * c + j] = -1;
        }
    }
}

}  // extern "C"
