// HighBitRecovery.map (data_process/process.py:726-751) on the device — the "high-bit recovery" of quantised dark frames on
// the real-data side of the noise path: every sample whose rounded DN value x lies in the LUT range [low, high) is re-drawn
// from the calibrated read-noise distribution conditioned on the quantisation cell,
//     x' = dist.ppf(cdf(x - 0.5) + U * (cdf(x + 0.5) - cdf(x - 0.5))),   U ~ U(0, 1),
// and the sub-DN remainder delta = value - round(value) is added back.  The LUT (cdf, range per integer x) is built on the
// host exactly as the reference does (scipy.stats); dist = Tukey-lambda(lam, loc = bias, scale = sigTL) if 'g' in the noise
// code, else N(bias, sigGs).  Arithmetic follows the reference's dtypes: the image stays float32, the quantile is float64 and
// is rounded to float32 on assignment.  U comes from Philox4x32-10 (53-bit, NumPy's construction) or from a caller array
// (replay of the reference's draws).
#include <cmath>
#include "abi_common.h"
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
        // scipy.stats.tukeylambda._ppf: boxcox(u, lam) - boxcox1p(-u, lam)
        if (fabs(a.lam) < 1e-19) q = log(u) - log1p(-u);
        else q = (expm1(a.lam * log(u)) - expm1(a.lam * log1p(-u))) / a.lam;
    } else {
        q = normcdfinv(u);
    }
    return q * a.scale + a.loc;
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
            r = (float)hbr_ppf(a.cdf[k] + u * a.range[k], a);
        }
        float o = __fadd_rn(r, delta);
        o = a.norm ? __fdiv_rn(o, a.span) : __fadd_rn(o, a.bl);
        a.out[i] = o;
    }
}

}  // namespace pnnp

using namespace pnnp;

extern "C" int pnnp_hbr_map(const float* in, float* out, size_t total, const double* cdf, const double* range, int low, int high,
                            int scale_in, int norm, float span, float bl, int dist_tukey, double lam, double loc, double scale,
                            const double* rand, uint64_t seed, uint64_t offset, uint64_t index0, double* rand_out, void* stream) {
    if (!in || !out || !cdf || !range || high < low) return fail("hbr_map: bad arguments");
    if (!total) return 0;
    HbrArgs a{in, out, total, cdf, range, low, high, scale_in, norm, span, bl, dist_tukey, lam, loc, scale, rand,
              philox_round_keys(seed), (uint32_t)offset, (uint32_t)(offset >> 32), index0, rand_out};
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    hbr_map_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
