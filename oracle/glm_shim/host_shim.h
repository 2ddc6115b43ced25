// TEST INFRASTRUCTURE ONLY (oracle/): force-included (-include) when the reference's .cu files are
// compiled by g++ for the host CPU.  Erases CUDA's function-space keywords, supplies the
// built-in index variables as thread-locals (one loop iteration plays one CUDA thread), and
// mirrors the NON-template global min/max overloads that CUDA injects for device code, so
// that unqualified scalar min/max in the reference resolve to fminf/fmaxf as they do under nvcc.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __align__(n) alignas(n)

struct orc_dim3 { unsigned int x, y, z; };
static thread_local orc_dim3 threadIdx, blockIdx, blockDim, gridDim;
static inline long long clock64() { return 0; }

static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline int min(int a, int b) { return b < a ? b : a; }
static inline int max(int a, int b) { return a < b ? b : a; }
