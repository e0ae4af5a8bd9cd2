// Per-sample arithmetic of the Bayer pack / unpack kernels (pack.cu), kept in a header so that the CPU suite can compile these
// very lines for the host (tests/emul/, test infrastructure only) and check them against the reference goldens without a GPU.
//   norm_one   raw2bayer's normalisation   utils/isp_ops.py:92-96
//   quant_one  bayer2raw's quantisation    utils/isp_ops.py:100-111
#pragma once
#include <cstdint>
#ifndef PNNP_HOST_EMUL
#include <cuda_runtime.h>
#endif

namespace pnnp {

__device__ __forceinline__ float norm_one(float v, double black, double wp, int norm, int clip) {
    if (norm) {
        const double d = __ddiv_rn(__dsub_rn((double)v, black), __dsub_rn(wp, black));
        v = (float)d;                       // clip in float64 then round == round then clip (0 and 1 are exact)
    }
    if (clip) v = fminf(fmaxf(v, 0.f), 1.f);
    return v;
}

// The same value without the float64 division subroutine (its ~30 instructions and the XU-pipe conversions around it, not the 6 bytes
// per sample, bounded the pack: 151 us per 64 crops against 61 us of HBM time).  The divisor wp - black is a per-plane constant, so
// with r = RN(1 / span) from the host, q = RN(x r), rem = x - span q (exact in one FMA), RN(q + rem r) == RN(x / span) (Markstein
// 1990; holds for every divisor whose significand is not all ones — the host checks and otherwise clears use_rcp).  v arrives as a
// double (a uint16 code or a float32 sample, both exact).
template <bool RCP>
__device__ __forceinline__ float norm_one_d(double v, double black, double wp, double span, double rcp, int norm, int clip) {
    float o;
    if (norm) {
        const double x = __dsub_rn(v, black);
        double d;
        if (RCP) {
            const double q = __dmul_rn(x, rcp);
            // a float32 sample is either below 3.5e38 or infinite / NaN: those stay what the division makes of them (q), the FMA
            // form would turn an infinity into NaN
            d = fabs(x) < 1e300 ? __fma_rn(__fma_rn(-span, q, x), rcp, q) : q;
        } else {
            d = __ddiv_rn(x, __dsub_rn(wp, black));
        }
        o = (float)d;
    } else {
        o = (float)v;
    }
    if (clip) o = fminf(fmaxf(o, 0.f), 1.f);
    return o;
}

__device__ __forceinline__ uint32_t quant_one(float v, float span, float bl) {
    v = fminf(fmaxf(v, 0.f), 1.f);
    v = __fadd_rn(__fmul_rn(v, span), bl);
    return (uint32_t)__float2uint_rz(v) & 0xFFFFu;       // numpy float32 -> uint16 cast truncates
}

}  // namespace pnnp
