// TEST INFRASTRUCTURE ONLY -- lets g++ compile radiation_ppo_b200/csrc/rs_env_impl.cuh as host code so that the
// kernel logic can be checked against the oracle in the (GPU-less) build container.  Never linked into the product.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__

struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
using std::max;
using std::min;

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline int __ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline void __threadfence() {}
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { uint32_t o = *p; *p |= v; return o; }
// round-down conversions used for the shortest-path lower bounds
static inline float __double2float_rd(double x) { float f = (float)x; if ((double)f > x) f = nextafterf(f, -INFINITY); return f; }
static inline float __int2float_rd(int x) { float f = (float)x; if ((double)f > (double)x) f = nextafterf(f, -INFINITY); return f; }
static inline float __fsqrt_rd(float v) { float f = sqrtf(v); if ((double)f * (double)f > (double)v) f = nextafterf(f, -INFINITY); return f; }
static inline float __double2float_ru(double x) { float f = (float)x; if ((double)f < x) f = nextafterf(f, INFINITY); return f; }
static inline float __fsub_ru(float a, float b) { return __double2float_ru((double)a - (double)b); }
static inline float __fmul_ru(float a, float b) { return __double2float_ru((double)a * (double)b); }
static inline float __fadd_rd(float a, float b) { return __double2float_rd((double)a + (double)b); }
static inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
