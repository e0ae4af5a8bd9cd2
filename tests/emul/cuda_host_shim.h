// Host shim for compiling pnnp_b200/csrc/noise_core.cuh with g++ (TEST INFRASTRUCTURE, tests/test_device_core_on_cpu.py).
// It lets the CPU suite run the DEVICE source of the samplers and of the deterministic tails against the oracle and the
// reference goldens.  Never part of the product: the library has no CPU path.
//
// Every CUDA intrinsic the core uses is an IEEE-754 round-to-nearest operation; with -ffp-contract=off the plain C++ operator
// is the same operation (x86-64 SSE2 arithmetic: no excess precision).  fmaf / fma are correctly rounded in glibc.  The three
// MUFU approximations (lg2 / ex2 / rsqrt) are replaced by libm: they only appear in the samplers, whose parity criterion is
// statistical.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
#define __restrict__

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

#include <cfenv>
static inline float __fadd_rd(float a, float b) {                 // add rounding towards -inf (FADD.RM)
    const int r = fegetround();
    fesetround(FE_DOWNWARD);
    volatile float va = a, vb = b;
    volatile float s = va + vb;
    fesetround(r);
    return s;
}
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline uint32_t __float2uint_rz(float a) { return a <= 0.f ? 0u : (a >= 4294967296.f ? 0xFFFFFFFFu : (uint32_t)a); }   // saturating, truncating
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __double2float_rn(double a) { return (float)a; }
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d; memcpy(&d, &b, 8); return d;
}

// ---- whole kernels on the host: a grid-stride kernel without shared memory / warp collectives runs exactly when its threads
// execute one after the other.  The driver (EMUL_LAUNCH) sets these and calls the kernel once per (block, thread).
#define __global__
#define __launch_bounds__(...)
struct EmulDim3 { unsigned x, y, z; };
static thread_local EmulDim3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#define EMUL_LAUNCH(grid, block, call)                                               \
    do {                                                                             \
        gridDim.x = (grid); blockDim.x = (block);                                    \
        for (unsigned b_ = 0; b_ < (unsigned)(grid); ++b_)                           \
            for (unsigned t_ = 0; t_ < (unsigned)(block); ++t_) { blockIdx.x = b_; threadIdx.x = t_; call; } \
    } while (0)
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, const T& v) { *p = v; }

// ---- bf16 (cuda_bf16.h): storage type + the three operations the layout kernels use
struct __nv_bfloat16 { uint16_t x; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat16 emul_f2bf(float f) {                 // round to nearest even, NaN kept quiet
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return __nv_bfloat16{(uint16_t)0x7FFF};
    u += 0x7FFFu + ((u >> 16) & 1u);
    return __nv_bfloat16{(uint16_t)(u >> 16)};
}
static inline float emul_bf2f(__nv_bfloat16 h) { return __uint_as_float((uint32_t)h.x << 16); }
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{emul_f2bf(a), emul_f2bf(b)}; }
static inline __nv_bfloat16 emul_bfmax(__nv_bfloat16 a, __nv_bfloat16 b) { return emul_bf2f(a) >= emul_bf2f(b) ? a : b; }
static inline __nv_bfloat162 __hmax2(__nv_bfloat162 a, __nv_bfloat162 b) { return __nv_bfloat162{emul_bfmax(a.x, b.x), emul_bfmax(a.y, b.y)}; }
static inline __nv_bfloat16 __float2bfloat16_rn(float f) { return emul_f2bf(f); }
static inline float __low2float(__nv_bfloat162 v) { return emul_bf2f(v.x); }
static inline float __high2float(__nv_bfloat162 v) { return emul_bf2f(v.y); }

// normcdfinv (CUDA math library, used by the HighBitRecovery map): Acklam's rational start, then two Halley steps on
// Phi(x) - p with libm's erfc -> accurate to a few ulp of double
static inline double normcdfinv(double p) {
    if (!(p > 0.0)) return p == 0.0 ? -INFINITY : NAN;
    if (!(p < 1.0)) return p == 1.0 ? INFINITY : NAN;
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
    double x;
    if (p < 0.02425) { const double q = sqrt(-2 * log(p)); x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1); }
    else if (p > 1 - 0.02425) { const double q = sqrt(-2 * log1p(-p)); x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1); }
    else { const double q = p - 0.5, r = q * q; x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1); }
    for (int it = 0; it < 2; ++it) {
        const double e = 0.5 * erfc(-x * 0.70710678118654752440) - p;
        const double u = e * 2.50662827463100050242 * exp(0.5 * x * x);
        x -= u / (1 + 0.5 * x * u);
    }
    return x;
}
