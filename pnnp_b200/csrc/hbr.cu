// HighBitRecovery.map (data_process/process.py:726-751) on the device — the "high-bit recovery" of quantised dark frames on
// the real-data side of the noise path: every sample whose rounded DN value x lies in the LUT range [low, high) is re-drawn
// from the calibrated read-noise distribution conditioned on the quantisation cell,
//     x' = dist.ppf(cdf(x - 0.5) + U * (cdf(x + 0.5) - cdf(x - 0.5))),   U ~ U(0, 1),
// and the sub-DN remainder delta = value - round(value) is added back.  The LUT (cdf, range per integer x) is built on the
// host exactly as the reference does (scipy.stats); dist = Tukey-lambda(lam, loc = bias, scale = sigTL) if 'g' in the noise
// code, else N(bias, sigGs).  Arithmetic follows the reference's dtypes: the image stays float32, the quantile is float64 and
// is rounded to float32 on assignment.  U comes from Philox4x32-10 (53-bit, NumPy's construction) or from a caller array
// (replay of the reference's draws).
#include <algorithm>
#include <cmath>
#include "abi_common.h"
#include "hbr_kernels.cuh"

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
