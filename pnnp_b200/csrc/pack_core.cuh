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

__device__ __forceinline__ uint32_t quant_one(float v, float span, float bl) {
    v = fminf(fmaxf(v, 0.f), 1.f);
    v = __fadd_rn(__fmul_rn(v, span), bl);
    return (uint32_t)__float2uint_rz(v) & 0xFFFFu;       // numpy float32 -> uint16 cast truncates
}

}  // namespace pnnp
