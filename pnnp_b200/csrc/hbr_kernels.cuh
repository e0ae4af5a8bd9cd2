// HighBitRecovery.map kernel (launcher and the C ABI: hbr.cu).  In a header so that the CPU suite can compile this very source for the
// host and run it thread by thread (tests/emul/; test infrastructure only).
#pragma once
#include <cmath>
#include "noise_core.cuh"

namespace pnnp {

struct HbrArgs {
    const float* in; float* out; size_t total;
    const double* cdf; const double* range; int low, high;
    int scale_in, norm; float span, bl;
    int dist_tukey; double lam, loc, scale;
    const double* rand; PhiloxKeys rk; uint32_t off_lo, off_hi; uint64_t index0;
    double* rand_out;
};

__device__ __forceinline__ double hbr_ppf(double u, const HbrArgs& a) {
    double q;
    if (a.dist_tukey) {
        // scipy.stats.tukeylambda._ppf = boxcox(u, lam) - boxcox1p(-u, lam); scipy/special/_boxcox.pxd divides each term by lam
        // before the subtraction: the same operations in the same order, none contracted
        if (fabs(a.lam) < 1e-19) q = __dsub_rn(log(u), log1p(-u));
        else q = __dsub_rn(__ddiv_rn(expm1(__dmul_rn(a.lam, log(u))), a.lam), __ddiv_rn(expm1(__dmul_rn(a.lam, log1p(-u))), a.lam));
    } else {
        q = normcdfinv(u);
    }
    return __dadd_rn(__dmul_rn(q, a.scale), a.loc);             // rv_continuous.ppf: _ppf(q) * scale + loc, two NumPy operations
}

__global__ void __launch_bounds__(256) hbr_map_kernel(const HbrArgs a) {
    const RngCtx rng{a.rk, a.off_lo, a.off_hi};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        const float df = a.scale_in ? __fmul_rn(a.in[i], a.span) : a.in[i];
        float r = rintf(df);                                   // np.round: half to even
        const float delta = __fsub_rn(df, r);
        double u;
        if (a.rand) u = a.rand[i];
        else {
            // element g uses words (2e, 2e+1) of the two Philox blocks of group g >> 1 ... one block per two elements
            const uint64_t g = a.index0 + i;
            const uint4 b = rng.block(g >> 1, 2u /* stream HBR */, 0u);
            const uint32_t w0 = (g & 1) ? b.z : b.x, w1 = (g & 1) ? b.w : b.y;
            u = ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) * (1.0 / 9007199254740992.0);   // NumPy's 53-bit double
        }
        if (a.rand_out) a.rand_out[i] = u;
        if (r >= (float)a.low && r < (float)a.high) {
            const int k = (int)r - a.low;
            r = (float)hbr_ppf(__dadd_rn(a.cdf[k], __dmul_rn(u, a.range[k])), a);
        }
        float o = (a.norm & 2) ? r : __fadd_rn(r, delta);       // HighBitRecovery(float=False) leaves the remainder out
        o = (a.norm & 1) ? __fdiv_rn(o, a.span) : __fadd_rn(o, a.bl);
        a.out[i] = o;
    }
}

}  // namespace pnnp
