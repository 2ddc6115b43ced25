// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for GLM 1.0.1, which the reference pins
// (README.md:3-4) but does not vendor (.gitignore:8 ignores cuda/includes/libraries/glm).
//
// This header lets the reference's own .cu files compile, unmodified and by path, for the
// host (g++) or for the device (nvcc).  It restates the *published semantics* of the GLM
// functions the mesh-generation path uses; each formula is the one in GLM 1.0.1's
// glm/detail/func_geometric.inl, func_common.inl and func_exponential.inl:
//   dot(a,b)      : tmp = a*b ; tmp.x + tmp.y + tmp.z          (compute_dot<vec<3>>)
//   length(v)     : sqrt(dot(v,v))
//   distance(a,b) : length(b - a)
//   inversesqrt(x): 1 / sqrt(x)
//   normalize(v)  : v * inversesqrt(dot(v,v))
//   cross(x,y)    : (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   mix(x,y,a)    : x*(1-a) + y*a
//   min(x,y)      : (y < x) ? y : x        max(x,y): (x < y) ? y : x
//   abs(x)        : x >= 0 ? x : -x
//   vec op scalar : component-wise, true division for '/'
// Scalar min/max/abs/mod/clamp are templates, as in GLM, so that the non-template CUDA /
// libm overloads win overload resolution exactly as they do in the reference's own build.
// Parity status: "unpinned" against real GLM (no copy is on disk); see DESIGN.md.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define GLMS_FN __host__ __device__ inline
#else
#define GLMS_FN inline
#endif

namespace glm {

typedef std::uint32_t u32;

template <int N, typename T> struct vec;

template <typename T> struct vec<2, T> {
    T x, y;
    GLMS_FN vec() : x(0), y(0) {}
    GLMS_FN explicit vec(T s) : x(s), y(s) {}
    template <typename A, typename B> GLMS_FN vec(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
    template <typename U> GLMS_FN vec(const vec<2, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
    template <typename U> GLMS_FN vec(const vec<3, U>& o);  // GLM_EXPLICIT is empty by default: implicit truncation (common.cu:121)
    GLMS_FN T& operator[](int i) { return i == 0 ? x : y; }
    GLMS_FN const T& operator[](int i) const { return i == 0 ? x : y; }
};

template <typename T> struct vec<3, T> {
    T x, y, z;
    GLMS_FN vec() : x(0), y(0), z(0) {}
    GLMS_FN explicit vec(T s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C>
    GLMS_FN vec(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
    template <typename U>
    GLMS_FN vec(const vec<3, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)) {}
    GLMS_FN T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    GLMS_FN const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    GLMS_FN vec& operator+=(const vec& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLMS_FN vec& operator-=(const vec& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    GLMS_FN vec& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
};

template <typename T> template <typename U>
GLMS_FN vec<2, T>::vec(const vec<3, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}

typedef vec<2, float> vec2;
typedef vec<3, float> vec3;
typedef vec<2, int> ivec2;
typedef vec<3, int> ivec3;
typedef vec<2, unsigned int> uvec2;

struct quat {
    float w, x, y, z;
    GLMS_FN quat() : w(1), x(0), y(0), z(0) {}
    GLMS_FN quat(float w_, float x_, float y_, float z_) : w(w_), x(x_), y(y_), z(z_) {}
};

#define GLMS_BINOP(OP)                                                                                         \
    template <typename T> GLMS_FN vec<3, T> operator OP(const vec<3, T>& a, const vec<3, T>& b) {              \
        return vec<3, T>(a.x OP b.x, a.y OP b.y, a.z OP b.z);                                                  \
    }                                                                                                          \
    template <typename T> GLMS_FN vec<3, T> operator OP(const vec<3, T>& a, T s) {                             \
        return vec<3, T>(a.x OP s, a.y OP s, a.z OP s);                                                        \
    }                                                                                                          \
    template <typename T> GLMS_FN vec<3, T> operator OP(T s, const vec<3, T>& a) {                             \
        return vec<3, T>(s OP a.x, s OP a.y, s OP a.z);                                                        \
    }                                                                                                          \
    template <typename T> GLMS_FN vec<2, T> operator OP(const vec<2, T>& a, const vec<2, T>& b) {              \
        return vec<2, T>(a.x OP b.x, a.y OP b.y);                                                              \
    }                                                                                                          \
    template <typename T> GLMS_FN vec<2, T> operator OP(const vec<2, T>& a, T s) {                             \
        return vec<2, T>(a.x OP s, a.y OP s);                                                                  \
    }                                                                                                          \
    template <typename T> GLMS_FN vec<2, T> operator OP(T s, const vec<2, T>& a) {                             \
        return vec<2, T>(s OP a.x, s OP a.y);                                                                  \
    }
GLMS_BINOP(+)
GLMS_BINOP(-)
GLMS_BINOP(*)
GLMS_BINOP(/)
#undef GLMS_BINOP

template <typename T> GLMS_FN vec<3, T> operator-(const vec<3, T>& a) { return vec<3, T>(-a.x, -a.y, -a.z); }
template <typename T> GLMS_FN vec<2, T> operator-(const vec<2, T>& a) { return vec<2, T>(-a.x, -a.y); }

// ---- geometric -------------------------------------------------------------------------
GLMS_FN float dot(const vec3& a, const vec3& b) { vec3 tmp(a * b); return tmp.x + tmp.y + tmp.z; }
GLMS_FN float dot(const vec2& a, const vec2& b) { vec2 tmp(a * b); return tmp.x + tmp.y; }
GLMS_FN float length(const vec3& v) { return sqrtf(dot(v, v)); }
GLMS_FN float length(const vec2& v) { return sqrtf(dot(v, v)); }
GLMS_FN float distance(const vec3& p0, const vec3& p1) { return length(p1 - p0); }
GLMS_FN float inversesqrt(float x) { return 1.0f / sqrtf(x); }
GLMS_FN vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
GLMS_FN vec3 cross(const vec3& x, const vec3& y) {
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}

// ---- common ----------------------------------------------------------------------------
GLMS_FN vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
template <typename T> GLMS_FN T mix(T x, T y, T a) { return x * (static_cast<T>(1) - a) + y * a; }
template <typename T> GLMS_FN T min(T x, T y) { return (y < x) ? y : x; }
template <typename T> GLMS_FN T max(T x, T y) { return (x < y) ? y : x; }
GLMS_FN vec3 min(const vec3& a, const vec3& b) { return vec3(min<float>(a.x, b.x), min<float>(a.y, b.y), min<float>(a.z, b.z)); }
GLMS_FN vec3 max(const vec3& a, const vec3& b) { return vec3(max<float>(a.x, b.x), max<float>(a.y, b.y), max<float>(a.z, b.z)); }
template <typename T> GLMS_FN T abs(T x) { return x >= static_cast<T>(0) ? x : -x; }
GLMS_FN vec3 abs(const vec3& a) { return vec3(abs<float>(a.x), abs<float>(a.y), abs<float>(a.z)); }
template <typename T> GLMS_FN T mod(T x, T y) { return x - y * floorf(x / y); }
template <typename T> GLMS_FN T clamp(T x, T lo, T hi) { return min<T>(max<T>(x, lo), hi); }
GLMS_FN vec2 floor(const vec2& a) { return vec2(floorf(a.x), floorf(a.y)); }
GLMS_FN vec2 round(const vec2& a) { return vec2(roundf(a.x), roundf(a.y)); }
GLMS_FN vec3 floor(const vec3& a) { return vec3(floorf(a.x), floorf(a.y), floorf(a.z)); }

// ---- the one matrix type the tone-mapper touches (column-major, like GLM) ----------------
struct mat3x3 {
    vec3 col[3];
    GLMS_FN mat3x3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2) {
        col[0] = vec3(x0, y0, z0); col[1] = vec3(x1, y1, z1); col[2] = vec3(x2, y2, z2);
    }
};
GLMS_FN vec3 operator*(const mat3x3& m, const vec3& v) { return m.col[0] * v.x + m.col[1] * v.y + m.col[2] * v.z; }

}  // namespace glm
