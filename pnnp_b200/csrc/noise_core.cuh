// Device-side core of the fused noise synthesis: Philox4x32-10, the samplers, and the
// deterministic arithmetic ("tail") shared by the Philox kernel and the replay kernel.
//
// Reference semantics: data_process/process.py:591-631 (generate_noisy_obs, NumPy chain) and
// :634-673 (generate_noisy_torch, float32 chain).  The cast-by-cast specification this file is
// written from is oracle/oracle_np.py::noisy_obs_tail_explicit / noisy_torch_tail.
//
// Every arithmetic step of the tails uses the *_rn intrinsics so that nvcc can never contract
// a mul+add into an FMA: the reference rounds after every NumPy / torch op.
#pragma once
#include <cstdint>
#ifndef PNNP_HOST_EMUL            // tests/emul/: the CPU suite compiles this file with g++ through a shim (test infrastructure only)
#include <cuda_runtime.h>
#endif
#include "../../include/pnnp_b200.h"

namespace pnnp {

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw — SC'11).  Counter layout used by this library:
//   key  = (seed_lo, seed_hi)
//   ctr0 = index_lo            ctr1 = index_hi[15:0] | sub << 16 | stream << 24
//   ctr2 = offset_lo           ctr3 = offset_hi
// index = global element index (stream ELEM) or global (crop, channel, row) index (stream ROW).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kStreamElem = 0u, kStreamRow = 1u;

// Round keys k + r*W for r = 0..9, precomputed on the host and passed in the kernel-parameter block: with the round loop fully
// unrolled every key is a constant-bank operand of the XOR, so a round is 2 IMAD.WIDE.U32 + 2 LOP3 and nothing else.
struct PhiloxKeys { uint32_t k[20]; };
inline PhiloxKeys philox_round_keys(uint64_t seed) {
    PhiloxKeys r;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) { r.k[2 * i] = k0; r.k[2 * i + 1] = k1; k0 += kPhiloxW0; k1 += kPhiloxW1; }
    return r;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& rk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ rk.k[2 * r], lo1, hi0 ^ c.w ^ rk.k[2 * r + 1], lo0);
    }
    return c;
}

// single-instruction special-function wrappers (the CUDA math-library forms add denormal-range fix-up code)
#ifndef PNNP_HOST_EMUL
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
inline float lg2_approx(float x) { return log2f(x); }
inline float ex2_approx(float x) { return exp2f(x); }
inline float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
#endif

struct RngCtx {
    const PhiloxKeys& rk;               // lives in the kernel-parameter (constant) bank
    uint32_t off_lo, off_hi;
    __device__ __forceinline__ uint4 block(uint64_t index, uint32_t stream, uint32_t sub) const {
        const uint32_t c1 = (uint32_t)((index >> 32) & 0xFFFFu) | (sub << 16) | (stream << 24);
        return philox4x32_10(make_uint4((uint32_t)index, c1, off_lo, off_hi), rk);
    }
};

// (0,1) with 24 random bits, exactly representable in float32
__device__ __forceinline__ float u01_24(uint32_t w) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-8f; }
// [0,1) with 24 random bits (torch.rand's lattice)
__device__ __forceinline__ float u01_24_closed0(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }
// (0,1] ∩ float32 from all 32 bits: relative precision kept near 0 (tails)
__device__ __forceinline__ float u01_32(uint32_t w) { return fmaf((float)w, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// ------------------------------------------------------------------------------------------
// Draw layout (all kernels).  Elements are grouped in fours by GLOBAL element index g; group G = g >> 2 owns the
// Philox blocks (G, stream ELEM, sub):
//   sub 0: word e            -> the shot-noise draw of element e = g & 3 (all 32 bits)
//   sub 1: word e = "mix"    -> read-noise cell k = mix >> 12 (20 bits) | quantisation draw = mix & 0xFFF (12 bits)
//   sub 2: word e            -> refinement of the read-noise draw, generated ONLY when k lies in the outer 256 cells of either
//                               tail (probability 2^-11): the low 12 bits of the read word come from it instead of the cell centre
// i.e. the read-noise sampler inverts the 32-bit word  w_read = k << 12 | (tail ? refine >> 20 : 0x800): 2^-20 resolution in
// the body (quantile step < 4e-4 sigma there), full 2^-32 resolution where the quantile function is steep.  The quantisation
// draw (U(-0.5, 0.5) DN, added to continuous read noise) has a 2^-12 DN lattice.  Two Philox blocks per four elements
// instead of three.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kReadTailCells = 256u;
__device__ __forceinline__ bool read_cell_is_tail(uint32_t k) { return (k - kReadTailCells) >= ((1u << 20) - 2u * kReadTailCells); }
__device__ __forceinline__ uint32_t read_word(uint32_t mix, uint32_t refine) {
    const uint32_t k = mix >> 12;
    return (mix & 0xFFFFF000u) | (read_cell_is_tail(k) ? (refine >> 20) : 0x800u);
}
// numpy chain: U(-0.5, 0.5) on the 12-bit lattice, exact in float64: (qbits + 0.5) / 4096 - 0.5 (built from the bit pattern of
// 1 + (qbits + 0.5) / 4096, no int->double conversion)
__device__ __forceinline__ double quant_draw_f64(uint32_t mix) {
    return __dadd_rn(__hiloint2double((int)(0x3FF00000u | ((mix & 0xFFFu) << 8) | 0x80u), 0), -1.5);
}
// torch chain: U[0, 1) on the 12-bit lattice
__device__ __forceinline__ float quant_draw_f32(uint32_t mix) { return (float)(mix & 0xFFFu) * 2.44140625e-4f; }

// ------------------------------------------------------------------------------------------
// Standard normal by inversion of ONE 32-bit word: z = Phi^-1((w + 0.5) / 2^32).
// The smaller tail probability p = (min(w, ~w) + 0.5) / 2^32 in (0, 0.5] is formed exactly, so both
// tails keep relative precision down to 2^-33.  With t = -log2(4 p (1-p)):
//   central (t < 8.25, p > ~8e-4): erfinv(x) = x * P7(t),  x = 1 - 2p
//   tail                          : erfinv(x) = Q6(sqrt(t))
// Coefficients: Chebyshev-node least-squares fits, tools/fit_normal_icdf.py (max |error| 1.3e-6 in z
// for the float32 evaluation, checked there against scipy.special.ndtri).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float normal_icdf(uint32_t w) {
    const bool lower = w < 0x80000000u;
    const uint32_t m = lower ? w : ~w;
    const float p = fmaf((float)m, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float t = -lg2_approx(4.0f * p * (1.0f - p));
    float e;
    if (t < 8.25f) {
        float q = 2.674857846e-08f;
        q = fmaf(q, t, -1.025935489e-06f);
        q = fmaf(q, t, 1.418562169e-05f);
        q = fmaf(q, t, -4.943624299e-05f);
        q = fmaf(q, t, -7.453467697e-04f);
        q = fmaf(q, t, 5.522758700e-03f);
        q = fmaf(q, t, 1.608277857e-01f);
        q = fmaf(q, t, 8.862264752e-01f);
        e = q * (1.0f - 2.0f * p);
    } else {
        const float s = sqrtf(t);
        float q = 7.718497727e-06f;
        q = fmaf(q, s, -2.602138266e-04f);
        q = fmaf(q, s, 3.719373606e-03f);
        q = fmaf(q, s, -2.902236022e-02f);
        q = fmaf(q, s, 1.309477687e-01f);
        q = fmaf(q, s, 5.167053342e-01f);
        q = fmaf(q, s, 1.426639557e-01f);
        e = q;
    }
    const float z = 1.4142135623730951f * e;
    return lower ? -z : z;
}

// ------------------------------------------------------------------------------------------
// Shot words.  The Poisson samplers read only the top 23 bits of their 32-bit word: the low 9 bits are NOT part of the draw
// (every sampler entry point replaces them by the cell centre 0x100), which leaves room for the specialised kernel to carry an
// element's position inside its 512-element unit through the sampler queue.  2^-23 is the float32 resolution of a uniform in
// [0.5, 1) anyway, and no conversion instruction is needed to form it: 0x3F800000 | bits is 1 + bits / 2^23.
// Conversions (I2F / F2I / F2F / FRND), POPC and MUFU all issue on the XU pipe, four lanes per SM sub-partition and clock:
// the r02 capture of the specialised kernel showed 14.7 of them per element = 38 % of the kernel's cycles on that pipe
// alone, so the samplers below avoid them (floor by an add rounding towards -inf into 2^23 + x, counts as integers).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kShotPosMask = 0x1FFu;
__device__ __forceinline__ uint32_t shot_word(uint32_t w) { return (w & ~kShotPosMask) | 0x100u; }
// floor(x) for 0 <= x < 2^23 as an integer and as a float, without F2I / FRND: the sum 2^23 + x has unit spacing
__device__ __forceinline__ float floor_bias23(float x) { return __fadd_rd(x, 8388608.0f); }
__device__ __forceinline__ int bias23_int(float t) { return (int)(__float_as_uint(t) & 0x7FFFFFu); }

// z = Phi^-1 of the 23-bit cell of a shot word (already through shot_word): sign bit + the 22 bits below it.  Same
// polynomials as normal_icdf; the central branch forms 2p = ((m >> 8) + 0.5) / 2^23 from the bit pattern, the tail branch
// (p < 8e-4) takes the integer conversion for its relative precision.
__device__ __forceinline__ float normal_icdf_cell23(uint32_t w) {
    const bool lower = w < 0x80000000u;
    const uint32_t m = lower ? w : ~w;
    const float p2 = __uint_as_float(0x3F800000u | (m >> 8)) - 0.99999994f;            // 2p in (0, 1), exact
    const float t = -lg2_approx(p2 * (2.0f - p2));
    float q = 2.674857846e-08f;
    q = fmaf(q, t, -1.025935489e-06f);
    q = fmaf(q, t, 1.418562169e-05f);
    q = fmaf(q, t, -4.943624299e-05f);
    q = fmaf(q, t, -7.453467697e-04f);
    q = fmaf(q, t, 5.522758700e-03f);
    q = fmaf(q, t, 1.608277857e-01f);
    q = fmaf(q, t, 8.862264752e-01f);
    float e = q * (1.0f - p2);
    if (t >= 8.25f) {                       // p < 8e-4: the straight-line body above is discarded (two independent entries interleave)
        const float p = fmaf((float)m, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        const float s = sqrtf(-lg2_approx(4.0f * p * (1.0f - p)));
        q = 7.718497727e-06f;
        q = fmaf(q, s, -2.602138266e-04f);
        q = fmaf(q, s, 3.719373606e-03f);
        q = fmaf(q, s, -2.902236022e-02f);
        q = fmaf(q, s, 1.309477687e-01f);
        q = fmaf(q, s, 5.167053342e-01f);
        q = fmaf(q, s, 1.426639557e-01f);
        e = q;
    }
    const float z = 1.4142135623730951f * e;
    return lower ? -z : z;
}

// ------------------------------------------------------------------------------------------
// Tukey-lambda quantile  Q(u) = (u^lam - (1-u)^lam) / lam   (lam -> 0: logit)
// u and 1-u are formed separately from the word so both tails keep relative precision.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float tukey_lambda_from_uv(float a, float b, float lam, float inv_lam);
// body cell (k + 0.5) / 2^20 and its complement, both exact in float32, without int->float conversions:
// 1 + (8k + 4) / 2^23 has the bit pattern 0x3F800000 | k << 3 | 4
__device__ __forceinline__ float tukey_lambda_ppf_body(uint32_t mix, float lam, float inv_lam) {
    const float a = __uint_as_float(0x3F800000u | ((mix >> 12) << 3) | 4u) - 1.0f;
    return tukey_lambda_from_uv(a, 1.0f - a, lam, inv_lam);
}
// power form only (|lam| >= 1e-3, which PNNP_CODE_UNIFORM_F64 promises): the specialised kernel's body cells
__device__ __forceinline__ float tukey_lambda_ppf_body_pow(uint32_t mix, float lam, float inv_lam) {
    const float a = __uint_as_float(0x3F800000u | ((mix >> 12) << 3) | 4u) - 1.0f;
    return (ex2_approx(lam * lg2_approx(a)) - ex2_approx(lam * lg2_approx(1.0f - a))) * inv_lam;
}
__device__ __forceinline__ float tukey_lambda_ppf(uint32_t w, float lam, float inv_lam) {
    if (!read_cell_is_tail(w >> 12)) return tukey_lambda_ppf_body(w, lam, inv_lam);
    return tukey_lambda_from_uv(u01_32(w), u01_32(~w), lam, inv_lam);
}
__device__ __forceinline__ float tukey_lambda_from_uv(float a, float b, float lam, float inv_lam) {
    const float la = lg2_approx(a), lb = lg2_approx(b);
    if (fabsf(lam) < 1e-3f) {
        const float xa = la * 0.6931471805599453f, xb = lb * 0.6931471805599453f;
        const float d1 = xa - xb, d2 = xa * xa - xb * xb, d3 = xa * xa * xa - xb * xb * xb;
        return d1 + lam * (0.5f * d2 + lam * 0.16666667f * d3);
    }
    return (ex2_approx(lam * la) - ex2_approx(lam * lb)) * inv_lam;
}

// ------------------------------------------------------------------------------------------
// Poisson sampler: inversion of ONE uniform word (no rejection loop, no per-lane retry divergence).
//   lam <  10 : exact inversion by sequential search of the CDF from k = 0 (the index k is
//               warp-uniform, so 1/k comes from the constant bank as a broadcast);
//   lam >= 10 : k = floor(Q(z)), z = Phi^-1(u), with the third-order Cornish-Fisher expansion of the
//               Poisson quantile including the continuity correction (Giles, "Algorithm 955:
//               approximation of the inverse Poisson CDF", ACM TOMS 42, 2016, eq. Q_N):
//                 Q = lam + sqrt(lam) z + (1/3 + z^2/6) + (-z/36 - z^3/72)/sqrt(lam)
//                       + (-8/405 + 7 z^2/810 + z^4/270)/lam
//               Its CDF differs from the exact Poisson CDF by at most 4.7e-5 at lam = 10, 3.9e-6 at 32,
//               < 1e-6 from 64 up (tests/test_poisson_model.py evaluates this against scipy on a grid);
//               a KS test needs ~1e9 samples at one rate to see it.  NumPy's legacy sampler (the
//               reference's) switches algorithm at the same lam = 10.
// ------------------------------------------------------------------------------------------
constexpr int kInvTab = 64;
__constant__ float c_inv_k[kInvTab] = {
    0.f, 1.f / 1, 1.f / 2, 1.f / 3, 1.f / 4, 1.f / 5, 1.f / 6, 1.f / 7, 1.f / 8, 1.f / 9, 1.f / 10, 1.f / 11, 1.f / 12,
    1.f / 13, 1.f / 14, 1.f / 15, 1.f / 16, 1.f / 17, 1.f / 18, 1.f / 19, 1.f / 20, 1.f / 21, 1.f / 22, 1.f / 23,
    1.f / 24, 1.f / 25, 1.f / 26, 1.f / 27, 1.f / 28, 1.f / 29, 1.f / 30, 1.f / 31, 1.f / 32, 1.f / 33, 1.f / 34,
    1.f / 35, 1.f / 36, 1.f / 37, 1.f / 38, 1.f / 39, 1.f / 40, 1.f / 41, 1.f / 42, 1.f / 43, 1.f / 44, 1.f / 45,
    1.f / 46, 1.f / 47, 1.f / 48, 1.f / 49, 1.f / 50, 1.f / 51, 1.f / 52, 1.f / 53, 1.f / 54, 1.f / 55, 1.f / 56,
    1.f / 57, 1.f / 58, 1.f / 59, 1.f / 60, 1.f / 61, 1.f / 62, 1.f / 63};
constexpr float kPoissonSwitch = 10.0f;

// w: a shot word that has been through shot_word().  Counts saturate at 2^23 - 1 (rates of millions of electrons per pixel:
// three orders of magnitude beyond any sensor's full well).
__device__ __forceinline__ int poisson_large_k(float lam, uint32_t w) {
    const float z = normal_icdf_cell23(w);
    const float rs = rsqrt_approx(lam), s = lam * rs, z2 = z * z;
    float x = fmaf(s, z, lam);
    x += fmaf(z2, 0.16666667f, 0.33333334f);
    x = fmaf(-z * fmaf(z2, 0.013888889f, 0.027777778f), rs, x);
    x = fmaf(fmaf(z2, fmaf(z2, 0.0037037036f, 0.008641975f), -0.019753087f), rs * rs, x);
    return bias23_int(floor_bias23(fminf(fmaxf(x, 0.f), 8388607.0f)));            // fmaxf also turns a NaN rate into 0
}

__device__ __forceinline__ float poisson_small(float lam, uint32_t w) {
    const float u = fminf(u01_32(w), 0.99999994f);
    float p = ex2_approx(-1.4426950408889634f * lam), F = p;
    int k = 0;
#pragma unroll 1
    while (u > F && k < kInvTab - 2) {          // two CDF terms per trip
        p *= lam * c_inv_k[k + 1];
        const float F1 = F + p;
        p *= lam * c_inv_k[k + 2];
        const float F2 = F1 + p;
        const bool stop1 = !(u > F1);
        k += stop1 ? 1 : 2;
        F = stop1 ? F1 : F2;
        if (stop1) break;
    }
    return (float)k;
}

// ------------------------------------------------------------------------------------------
// lam < 10, table form of the same exact inversion (what the kernels use; poisson_small above stays as the fallback for
// counts beyond the table and as the specification).  The sequential search costs every lane of a warp the trip count of
// its slowest lane; this form has a fixed, short instruction sequence instead:
//   * T[i][k] = P(Poisson(i/16) <= k), i = 0..160, k = 0..31: float32 rounded from a float64 evaluation on the host, copied
//     to shared memory by every CTA (22.5 KB; five leading zeros per row stand for k < 0).
//   * For lam = i/16 + delta (0 <= delta < 1/16), Poisson(lam) = Poisson(i/16) + Poisson(delta) gives the CDF at lam
//     EXACTLY as a short convolution of row i:  F(k; lam) = e^-delta * sum_j delta^j/j! * T[i][k-j];  terms j <= 4 are kept
//     (truncation <= delta^5/120 = 8e-9, below the float32 resolution of the table).
//   * T[i][k] >= F(k; lam), so a 5-probe binary search of row i for the first entry >= u gives a lower bound k0 of the
//     answer; the convolved CDF is then checked at k0, k0+1, ... (the first check succeeds with probability ~1 - delta/2).
// ------------------------------------------------------------------------------------------
constexpr int kPoisRows = 161, kPoisCols = 32, kPoisPad = 5, kPoisStride = kPoisCols + kPoisPad;   // 37 floats per row
constexpr int kPoisTableFloats = kPoisRows * kPoisStride;

// w: a shot word (its low 9 bits are ignored: u has 23 bits).  Two steps, so that a caller can run the straight-line first
// step of several entries side by side (their shared-memory probes are a chain of five dependent loads each) before the
// data-dependent second step.
struct PoisSmall { const float* row; float delta, c2, c3, c4, ue, t1, t2, t3, t4; int k; };
__device__ __forceinline__ PoisSmall poisson_small_search(float lam, uint32_t w, const float* __restrict__ T) {
    PoisSmall s;
    lam = fminf(fmaxf(lam, 0.f), 9.999999f);                                   // also keeps the row index inside the table
    const float u = __uint_as_float(0x3F800000u | (w >> 9)) - 0.99999994f;     // ((w >> 9) + 0.5) / 2^23 in (0, 1), exact
    const float ti = floor_bias23(lam * 16.0f);                                // 2^23 + floor(16 lam)
    s.delta = fmaf(-0.0625f, ti - 8388608.0f, lam);                            // exact
    s.row = T + bias23_int(ti) * kPoisStride + kPoisPad;
    int k = 0;
#pragma unroll
    for (int st = 16; st >= 1; st >>= 1) k += (s.row[k + st - 1] < u) ? st : 0;   // entries 0..30 that are < u
    s.k = k;
    s.c2 = 0.5f * s.delta * s.delta; s.c3 = s.c2 * s.delta * 0.33333334f; s.c4 = s.c3 * s.delta * 0.25f;
    // u * e^delta (delta < 1/16: degree-5 Taylor, relative error 1e-10)
    const float d = s.delta;
    s.ue = u * fmaf(d, fmaf(d, fmaf(d, fmaf(d, fmaf(d, 8.3333333e-3f, 4.1666667e-2f), 0.16666667f), 0.5f), 1.0f), 1.0f);
    s.t1 = s.row[k - 1]; s.t2 = s.row[k - 2]; s.t3 = s.row[k - 3]; s.t4 = s.row[k - 4];
    return s;
}
__device__ __forceinline__ int poisson_small_finish(PoisSmall& s, float lam, uint32_t w) {
#pragma unroll 1
    while (true) {
        if (s.k >= kPoisCols) return (int)poisson_small(fminf(fmaxf(lam, 0.f), 9.999999f), shot_word(w));   // probability < 1e-8: sequential search
        const float t0 = s.row[s.k];
        if (fmaf(s.c4, s.t4, fmaf(s.c3, s.t3, fmaf(s.c2, s.t2, fmaf(s.delta, s.t1, t0)))) >= s.ue) break;
        s.t4 = s.t3; s.t3 = s.t2; s.t2 = s.t1; s.t1 = t0; ++s.k;
    }
    return s.k;
}
__device__ __forceinline__ int poisson_small_table_k(float lam, uint32_t w, const float* __restrict__ T) {
    PoisSmall s = poisson_small_search(lam, w, T);
    return poisson_small_finish(s, lam, w);
}

// T: the shared-memory copy of the table above
__device__ __forceinline__ float poisson_sample(float lam, uint32_t w, const float* __restrict__ T) {
    if (!(lam > 0.f)) return 0.f;
    w = shot_word(w);
    return (float)(lam < kPoissonSwitch ? poisson_small_table_k(lam, w, T) : poisson_large_k(lam, w));
}

// ------------------------------------------------------------------------------------------
// Per-crop parameters as the kernels hold them (one table row, read through L1 by every lane)
// ------------------------------------------------------------------------------------------
struct RowP {
    double K, sigTL, sigGs, sigR, lam, q, ratio, span, lo;
    uint32_t flags;
};
__device__ __forceinline__ RowP load_row_params(const pnnp_noise_params* t) {
    RowP p;
    p.K = t->K; p.sigTL = t->sigTL; p.sigGs = t->sigGs; p.sigR = t->sigR; p.lam = t->lam; p.q = t->q;
    p.ratio = t->ratio; p.span = t->span; p.lo = t->clip_lo; p.flags = t->flags;
    return p;
}

// Scale-in (process.py:593-595) and the Poisson rate / Gaussian-shot scale the reference forms.
struct ScaleIn { float ysc32; double ysc64; };
__device__ __forceinline__ ScaleIn scale_in_numpy(float y, const RowP& p) {
    ScaleIn s;
    const float y32 = __fmul_rn(y, (float)p.span);
    if (p.flags & PNNP_F_RATIO64) { s.ysc64 = __ddiv_rn((double)y32, p.ratio); s.ysc32 = (float)s.ysc64; }
    else { s.ysc32 = __fdiv_rn(y32, (float)p.ratio); s.ysc64 = (double)s.ysc32; }
    return s;
}
__device__ __forceinline__ float poisson_rate_numpy(const ScaleIn& s, const RowP& p) {
    // 1.0*y/K in K's precision (float64 if either side is float64); our sampler consumes float32
    if (p.flags & (PNNP_F_K64 | PNNP_F_RATIO64)) return (float)(s.ysc64 / p.K);
    return __fdiv_rn(s.ysc32, (float)p.K);
}

// np.clip(z, lo, hi) on float64 as two compare+select pairs (fmin/fmax carry NaN-quieting logic that costs ~10 instructions
// each in SASS; np.clip propagates NaN, which a plain select does too)
__device__ __forceinline__ double clip_f64(double z, double lo, double hi) {
    z = z < lo ? lo : z;
    return z > hi ? hi : z;
}

// ------------------------------------------------------------------------------------------
// Tail, NumPy chain.  d_shot = Poisson count (code P) or N(0,1) draw; d_read in DN;
// d_rowz = N(0,1) row draw; d_q = U(-.5,.5) in DN (float64); bias_c = per-channel bias.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float tail_numpy(float y, const RowP& p, uint32_t code, bool ori, bool clip01,
                                            float d_shot, float d_read, float d_rowz, double d_q, double bias_c) {
    const bool k64 = p.flags & PNNP_F_K64, r64 = p.flags & PNNP_F_RATIO64, s64 = p.flags & PNNP_F_SIG64;
    const float span32 = (float)p.span;
    const ScaleIn s = scale_in_numpy(y, p);
    bool a64;
    float a32 = 0.f;
    double A = 0.0;
    if (code & PNNP_CODE_P) {
        if (k64) { A = __dmul_rn((double)d_shot, p.K); a64 = true; }
        else     { a32 = __fmul_rn(d_shot, (float)p.K); a64 = false; }
    } else if (k64 || r64) {
        const double yy = s.ysc64;
        const double t = __dsqrt_rn(fmax(__ddiv_rn(yy, p.K), 1e-10));
        A = __dadd_rn(yy, __dmul_rn(__dmul_rn((double)d_shot, t), p.K));
        a64 = true;
    } else {
        const float K32 = (float)p.K;
        const float t = __fsqrt_rn(fmaxf(__fdiv_rn(s.ysc32, K32), 1e-10f));
        a32 = __fadd_rn(s.ysc32, __fmul_rn(__fmul_rn(d_shot, t), K32));
        a64 = false;
    }
    if (!(code & PNNP_CODE_B)) {
        if (a64) A = __dadd_rn(A, (double)d_read); else a32 = __fadd_rn(a32, d_read);
        if (code & PNNP_CODE_R) {
            if (s64) {
                if (!a64) { A = (double)a32; a64 = true; }
                A = __dadd_rn(A, __dmul_rn((double)d_rowz, p.sigR));
            } else {
                const float row32 = __fmul_rn(d_rowz, (float)p.sigR);
                if (a64) A = __dadd_rn(A, (double)row32); else a32 = __fadd_rn(a32, row32);
            }
        }
        if (code & PNNP_CODE_Q) {
            if (!a64) { A = (double)a32; a64 = true; }
            A = __dadd_rn(A, d_q);
        }
        if (code & PNNP_CODE_D) {
            if (!a64) { A = (double)a32; a64 = true; }
            A = __dadd_rn(A, bias_c);
        }
    }
    if (a64) {
        double z = __ddiv_rn(A, p.span);
        z = clip01 ? clip_f64(z, 0.0, 1.0) : clip_f64(z, p.lo, 1.0);
        if (!ori) z = __dmul_rn(z, p.ratio);
        return (float)z;
    }
    float z = __fdiv_rn(a32, span32);
    z = clip01 ? fminf(fmaxf(z, 0.f), 1.f) : fminf(fmaxf(z, (float)p.lo), 1.f);
    if (!ori) {
        if (r64) return (float)__dmul_rn((double)z, p.ratio);
        z = __fmul_rn(z, (float)p.ratio);
    }
    return z;
}

// ------------------------------------------------------------------------------------------
// Correctly rounded division by a per-crop constant without the division subroutine (Markstein 1990):
// with r = RN(1/b), q = RN(a r), rem = a - b q (exact in one FMA), RN(q + rem r) == RN(a / b).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float div_rn_by_const(float a, float b, float r) {
    const float q = __fmul_rn(a, r);
    return __fmaf_rn(__fmaf_rn(-b, q, a), r, q);
}
__device__ __forceinline__ double div_rn_by_const(double a, double b, double r) {
    const double q = __dmul_rn(a, r);
    return __fma_rn(__fma_rn(-b, q, a), r, q);
}

// Per-crop constants of the specialised ("fast") NumPy-chain path: noise_code 'p','g','r','q' only,
// K / sigR np.float64 and ratio a python float (what sample_params returns), ori=False, clip=False.
struct FastC {
    float span32, ratio32, rratio32, invK32, sigTL32, lam_tl, inv_lam_tl, out_lo, out_hi;
    double K, span, rspan, ratio, sigR;
};
__device__ __forceinline__ FastC fast_constants(const pnnp_noise_params* t, float post_lo, float post_hi) {
    FastC f;
    f.K = t->K; f.span = t->span; f.ratio = t->ratio; f.sigR = t->sigR;
    f.rspan = __drcp_rn(f.span);
    f.span32 = (float)f.span; f.ratio32 = (float)f.ratio; f.rratio32 = __frcp_rn(f.ratio32);
    f.invK32 = (float)(1.0 / f.K); f.sigTL32 = (float)t->sigTL; f.lam_tl = (float)t->lam;
    f.inv_lam_tl = f.lam_tl != 0.f ? 1.0f / f.lam_tl : 0.f;
    // np.clip(z, lo, 1) * ratio -> float32 -> post-clip is a composition of monotone maps of z, so both clips commute with the
    // multiplication and the final rounding: float32(clip(z, lo, 1) * ratio) == clamp(float32(z * ratio), float32(lo * ratio),
    // float32(1 * ratio)), and a clamp followed by the post-clip clamp is ONE clamp whose bounds are the post-clipped bounds.
    // Two FMNMX per element instead of two float64 compare + select pairs and two FMNMX.  (A negative ratio swaps the bounds.)
    const float a = (float)__dmul_rn(t->clip_lo, f.ratio), b = (float)__dmul_rn(1.0, f.ratio);
    f.out_lo = fminf(fmaxf(fminf(a, b), post_lo), post_hi);
    f.out_hi = fminf(fmaxf(fmaxf(a, b), post_lo), post_hi);
    return f;
}
// Poisson rate of a clean pixel (scale-in + 1.0*y/K rounded to float32 for the samplers)
__device__ __forceinline__ float fast_rate(const FastC& f, float y) {
    return div_rn_by_const(__fmul_rn(y, f.span32), f.ratio32, f.rratio32) * f.invK32;
}
// count -> float64 exactly, without the conversion instruction: 2^52 + cnt has unit spacing (0 <= cnt < 2^31)
__device__ __forceinline__ double count_to_f64(int cnt) { return __dadd_rn(__hiloint2double(0x43300000, cnt), -4503599627370496.0); }
__device__ __forceinline__ float fast_tail(const FastC& f, int cnt, float d_read, double row64, double d_q) {
    double A = __dmul_rn(count_to_f64(cnt), f.K);
    A = __dadd_rn(A, (double)d_read);
    A = __dadd_rn(A, row64);
    A = __dadd_rn(A, d_q);
    const float v = (float)__dmul_rn(div_rn_by_const(A, f.span, f.rspan), f.ratio);
    return fminf(fmaxf(v, f.out_lo), f.out_hi);
}

// ------------------------------------------------------------------------------------------
// Tail, torch chain (float32 throughout; d_q = U[0,1) draw, q = (u-0.5)*q*(wp-bl)).
// 'b' zeroes only the read term there (process.py:652-661).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float scale_in_torch(float y, const RowP& p) {
    return __fdiv_rn(__fmul_rn(y, (float)p.span), (float)p.ratio);
}
__device__ __forceinline__ float tail_torch(const RowP& p, uint32_t code, bool ori, bool clip01, float d_shot,
                                            float d_read, float d_rowz, float d_qu) {
    const float span32 = (float)p.span, K32 = (float)p.K, ratio32 = (float)p.ratio;
    float acc = __fmul_rn(d_shot, K32);
    acc = __fadd_rn(acc, (code & PNNP_CODE_B) ? 0.f : d_read);
    if (code & PNNP_CODE_R) acc = __fadd_rn(acc, __fmul_rn(d_rowz, (float)p.sigR));
    if (code & PNNP_CODE_Q)
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(__fsub_rn(d_qu, 0.5f), (float)p.q), span32));
    float z = __fdiv_rn(acc, span32);
    z = clip01 ? fminf(fmaxf(z, 0.f), 1.f) : fminf(fmaxf(z, (float)p.lo), 1.f);
    if (!ori) z = __fmul_rn(z, ratio32);
    return z;
}

}  // namespace pnnp
