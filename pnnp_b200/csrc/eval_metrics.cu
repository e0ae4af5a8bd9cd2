// Eval boundary on the device (E1/E2): trainer_SID.py:231-248, data_process/__init__.py:162-175,
// utils/visualization.py:9-31.
//
//   dn' = clamp(dn * ratio?, 0, 1)                                     (trainer_SID.py:231-235)
//   IlluminanceCorrect: gain = <dn', hr> / <dn', dn'> over hr != 1     (data_process/__init__.py:162-175)
//   tensor2im: x255, clip [0, 255], no rounding                        (utils/visualization.py:9-24)
//   PSNR  = 10 log10(255^2 / MSE)          (skimage.metrics.peak_signal_noise_ratio, float64 MSE)
//   SSIM  = skimage.metrics.structural_similarity defaults: 7x7 uniform window, K1 .01, K2 .03,
//           sample covariance, 3-pixel border cropped, mean over pixels then channels.
// Instead of copying two 48.5 MB frames to the host and running skimage there, the kernels leave
// per-frame partial sums (a few doubles) that the host — or one NCCL all-reduce — finishes.
//
// sums layout per frame (double): [0] num  [1] den  [2] sum sq err (x255 domain)
//                                 [3 .. 3+c) per-channel sum of the SSIM map over valid centres
#include <cstdlib>
#include "abi_common.h"
#include "eval_kernels.cuh"
#include "../../include/pnnp_b200.h"

using namespace pnnp;

extern "C" int pnnp_eval_epilogue(const float* dn, const float* hr, int n, int c, int h, int w, float scale,
                                  int brightness_correct, double* sums, void* stream) {
    if (!dn || !hr || !sums) return fail("eval_epilogue: null pointer");
    if (n <= 0 || c <= 0 || h < kSsimWin || w < kSsimWin) return fail("eval_epilogue: frame smaller than the 7x7 SSIM window");
    cudaStream_t st = (cudaStream_t)stream;
    const int stride = 3 + c;
    PNNP_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)n * stride, st));
    const size_t per_frame = (size_t)c * h * w;
    if (brightness_correct) {
        dim3 g((unsigned)std::min<size_t>((per_frame + 255) / 256, 592), (unsigned)n);
        illum_dots_kernel<<<g, 256, 0, st>>>(dn, hr, per_frame, scale, sums, stride);
        count_launch();
    }
    if (variant_on("PNNP_SSIM_V2")) {          // separable form (ssim_core.cuh), the default since r02
        static bool attr_done = false;
        if (!attr_done) {
            PNNP_CUDA(cudaFuncSetAttribute(ssim_mse_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Ssim2Tile)));
            attr_done = true;
        }
        Ssim2Args a{dn, hr, c, h, w, scale, 1.0f, brightness_correct};
        dim3 gv((w + kS2TileX - 1) / kS2TileX, ((h + kS2TileY - 1) / kS2TileY + kS2TilesPerCta - 1) / kS2TilesPerCta, n * c);
        ssim_mse_v2_kernel<<<gv, kS2Threads, sizeof(Ssim2Tile), st>>>(a, sums, stride);
        count_launch();
        PNNP_CUDA(cudaGetLastError());
        return 0;
    }
    dim3 g2((w + kTileX - 1) / kTileX, (h + kTileY - 1) / kTileY, n * c);
    ssim_mse_kernel<<<g2, 256, 0, st>>>(dn, hr, c, h, w, scale, brightness_correct, sums, stride);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
