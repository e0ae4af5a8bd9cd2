// Whole CTAs of the warp-collective kernels compiled for the HOST and run on the lock-step fibre emulator (simt_host.h) — TEST
// INFRASTRUCTURE ONLY (tests/test_device_simt_on_cpu.py).  Same kernel source as the GPU build (csrc/noise_kernels.cuh).
#define PNNP_HOST_EMUL 1
#include "cuda_host_shim.h"
#include "simt_host.h"
#include "../../pnnp_b200/csrc/noise_kernels.cuh"
#include "../../pnnp_b200/csrc/train_kernels.cuh"
#include "../../pnnp_b200/csrc/eval_kernels.cuh"
#include "../../pnnp_b200/csrc/hbr_kernels.cuh"

namespace pnnp {                                     // the kernels' dynamic shared memory (`extern __shared__` arrays)
alignas(16) uint8_t s_fast[kFastSmemBytes];
float s_acc[4 * 64 + 4 + 64], s_b[1024], s_b2[2048], s_pool_bias[2048];
alignas(16) uint8_t s2_raw[sizeof(Ssim2Tile)];
}

using namespace pnnp;

extern "C" {

// kernel: 0 = generic, 128-bit path; 1 = generic, scalar path; 2 = specialised ("fast") kernel.  blocks is a free parameter:
// results must not depend on it.  debug != 0 records the draws (d_* may be null individually).
int emul_noise_synth(const float* clean, float* noisy, const pnnp_noise_params* table, int n, int c, int h, int w, uint32_t code,
                     int chain, int ori, int clip, float post_lo, float post_hi, uint64_t seed, uint64_t offset, uint64_t crop_id0,
                     int kernel, int debug, float* d_shot, float* d_read, float* d_rowz, double* d_q, int blocks) {
    static bool table_done = false;
    if (!table_done) { build_poisson_table(g_pois_table); table_done = true; }
    SynthArgs a{clean, noisy, table, n, c, h, w, code & 0x3Fu, ori, clip, post_lo, post_hi, seed, offset, crop_id0,
                philox_round_keys(seed), d_shot, d_read, d_rowz, d_q};
    if (kernel == 2) {
        if (w % 4 || chain != PNNP_CHAIN_NUMPY || a.code != (PNNP_CODE_P | PNNP_CODE_G | PNNP_CODE_R | PNNP_CODE_Q) || ori || clip) return 1;
        if (debug) SIMT_LAUNCH(blocks, kFastThreads, (noise_synth_fast_kernel<true>(a)));
        else SIMT_LAUNCH(blocks, kFastThreads, (noise_synth_fast_kernel<false>(a)));
        return 0;
    }
    if (kernel == 0 && (w % 4 || (crop_id0 * (uint64_t)c * h * w) % 4)) return 1;
#define EMUL_K(CH, DBG, V) SIMT_LAUNCH(blocks, kThreads, (noise_synth_kernel<CH, DBG, V>(a)))
    if (chain == PNNP_CHAIN_NUMPY) {
        if (kernel == 0) { if (debug) EMUL_K(PNNP_CHAIN_NUMPY, true, 4); else EMUL_K(PNNP_CHAIN_NUMPY, false, 4); }
        else             { if (debug) EMUL_K(PNNP_CHAIN_NUMPY, true, 1); else EMUL_K(PNNP_CHAIN_NUMPY, false, 1); }
    } else {
        if (kernel == 0) { if (debug) EMUL_K(PNNP_CHAIN_TORCH, true, 4); else EMUL_K(PNNP_CHAIN_TORCH, false, 4); }
        else             { if (debug) EMUL_K(PNNP_CHAIN_TORCH, true, 1); else EMUL_K(PNNP_CHAIN_TORCH, false, 1); }
    }
#undef EMUL_K
    return 0;
}

// replay kernel (pnnp_noise_synth_replay): the tails on caller-supplied draws
int emul_noise_replay(const float* clean, float* noisy, const pnnp_noise_params* table, int n, int c, int h, int w, uint32_t code, int chain,
                      int ori, int clip, float post_lo, float post_hi, const float* d_shot, const float* d_read, const float* d_rowz,
                      const double* d_q, int blocks) {
    SynthArgs a{clean, noisy, table, n, c, h, w, code, ori, clip, post_lo, post_hi, 0, 0, 0, PhiloxKeys{},
                const_cast<float*>(d_shot), const_cast<float*>(d_read), const_cast<float*>(d_rowz), const_cast<double*>(d_q)};
    if (chain == PNNP_CHAIN_NUMPY) SIMT_LAUNCH(blocks, 256, (noise_replay_kernel<PNNP_CHAIN_NUMPY>(a)));
    else SIMT_LAUNCH(blocks, 256, (noise_replay_kernel<PNNP_CHAIN_TORCH>(a)));
    return 0;
}

// ---- training-step helpers (csrc/train_kernels.cuh); `blocks` is free wherever the launcher's choice is not part of the contract
int emul_l1_loss(const float* pred, const float* hr, float* gpred, size_t total, double* loss_sum, int blocks) {
    *loss_sum = 0.0;
    SIMT_LAUNCH(blocks, 256, (l1_loss_kernel(pred, hr, gpred, total, 1.0f / (float)total, loss_sum)));
    return 0;
}
int emul_head_bwd(const float* gpred, const uint16_t* act, const float* W, uint16_t* gact, float* dW, float* db, float* dbias_prev,
                  int n, int h, int w, int cin, int co, int act_kind, int blocks) {
    if (co < 1 || co > 4 || (cin != 8 && cin != 16 && cin != 32 && cin != 64)) return 1;
    SIMT_LAUNCH(blocks, 256, (head_bwd_kernel(gpred, reinterpret_cast<const __nv_bfloat16*>(act), W, reinterpret_cast<__nv_bfloat16*>(gact),
                                              dW, db, dbias_prev, n, h, w, cin, co, act_kind)));
    return 0;
}
int emul_act_bwd_bias(uint16_t* g, const uint16_t* out, float* dbias, size_t pixels, int c, int act_kind, int blocks) {
    if ((c % 8) || c > 1024 || ((size_t)blocks * 256) % (size_t)(c / 8)) return 1;       // the launcher keeps grid * 256 a multiple of c / 8
    SIMT_LAUNCH(blocks, 256, (act_bwd_bias_kernel(reinterpret_cast<__nv_bfloat16*>(g), reinterpret_cast<const __nv_bfloat16*>(out), dbias,
                                                  pixels, c, act_kind)));
    return 0;
}
int emul_act_bwd_bias_v2_kernel(uint16_t* g, const uint16_t* out, float* dbias, uint32_t items, int c, int act_kind, int blocks) {
    if (c > 2048 || ((c / 8) & (c / 8 - 1))) return 1;
    const ActBwd2Args a{g, out, items, c, act_kind};
    SIMT_LAUNCH(blocks, kAb2Threads, (act_bwd_bias_v2_kernel(a, dbias)));
    return 0;
}
int emul_maxpool_bwd(const uint16_t* gp, const uint16_t* cfull, const uint16_t* gskip, uint16_t* gc, int n, int h, int w, int c,
                     int act_kind, int blocks) {
    if ((h & 1) || (w & 1) || (c & 7)) return 1;
    SIMT_LAUNCH(blocks, 256, (maxpool_bwd_kernel<false>(reinterpret_cast<const __nv_bfloat16*>(gp), reinterpret_cast<const __nv_bfloat16*>(cfull),
                                                        reinterpret_cast<const __nv_bfloat16*>(gskip), reinterpret_cast<__nv_bfloat16*>(gc), nullptr,
                                                        n, h, w, c, act_kind)));
    return 0;
}
int emul_maxpool_bwd_bias(const uint16_t* gp, const uint16_t* cfull, const uint16_t* gskip, uint16_t* gc, float* dbias, int n, int h, int w,
                          int c, int act_kind, int blocks) {
    if ((h & 1) || (w & 1) || (c & 7) || (256 % (c / 8))) return 1;
    SIMT_LAUNCH(blocks, 256, (maxpool_bwd_kernel<true>(reinterpret_cast<const __nv_bfloat16*>(gp), reinterpret_cast<const __nv_bfloat16*>(cfull),
                                                       reinterpret_cast<const __nv_bfloat16*>(gskip), reinterpret_cast<__nv_bfloat16*>(gc), dbias,
                                                       n, h, w, c, act_kind)));
    return 0;
}
int emul_adam(float* p, const float* g, float* m, float* v, size_t total, float lr, float b1, float b2, float eps, int step, float gscale,
              int on_device_state, int blocks) {
    if (on_device_state) {                       // pnnp_adam_step_dev: learning rate and step count read by the kernel
        float state[2] = {lr, (float)(step - 1)};
        adam_tick_kernel(state);
        SIMT_LAUNCH(blocks, 256, (adam_dev_kernel(p, g, m, v, total, state, b1, b2, eps, gscale)));
    } else {
        const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
        SIMT_LAUNCH(blocks, 256, (adam_kernel(p, g, m, v, total, lr, b1, b2, eps, bc1, bc2, gscale)));
    }
    return 0;
}

// pnnp_adam_step_dev as the training step calls it: learning rate and step count live in `state` (advanced by the tick kernel)
int emul_adam_dev(float* p, const float* g, float* m, float* v, size_t total, float* state, float b1, float b2, float eps, float gscale, int blocks) {
    adam_tick_kernel(state);
    SIMT_LAUNCH(blocks, 256, (adam_dev_kernel(p, g, m, v, total, state, b1, b2, eps, gscale)));
    return 0;
}

// ---- eval epilogue (csrc/eval_kernels.cuh), as pnnp_eval_epilogue launches it; dots_blocks is free
int emul_eval_epilogue(const float* dn, const float* hr, int n, int c, int h, int w, float scale, int brightness_correct, double* sums,
                       int v2, int dots_blocks) {
    if (n <= 0 || c <= 0 || h < kSsimWin || w < kSsimWin) return 1;
    const int stride = 3 + c;
    for (int i = 0; i < n * stride; ++i) sums[i] = 0.0;
    const size_t per_frame = (size_t)c * h * w;
    if (brightness_correct) SIMT_LAUNCH3(dots_blocks, n, 1, 256, (illum_dots_kernel(dn, hr, per_frame, scale, sums, stride)));
    if (v2) {
        const Ssim2Args a{dn, hr, c, h, w, scale, 1.0f, brightness_correct};
        SIMT_LAUNCH3((w + kS2TileX - 1) / kS2TileX, ((h + kS2TileY - 1) / kS2TileY + kS2TilesPerCta - 1) / kS2TilesPerCta, n * c, kS2Threads,
                     (ssim_mse_v2_kernel(a, sums, stride)));
    } else {
        SIMT_LAUNCH3((w + kTileX - 1) / kTileX, (h + kTileY - 1) / kTileY, n * c, 256,
                     (ssim_mse_kernel(dn, hr, c, h, w, scale, brightness_correct, sums, stride)));
    }
    return 0;
}

// ---- HighBitRecovery map (csrc/hbr_kernels.cuh), arguments of pnnp_hbr_map
int emul_hbr_map(const float* in, float* out, size_t total, const double* cdf, const double* range, int low, int high, int scale_in,
                 int norm, float span, float bl, int dist_tukey, double lam, double loc, double scale, const double* rand, uint64_t seed,
                 uint64_t offset, uint64_t index0, double* rand_out, int blocks) {
    if (high < low) return 1;
    const HbrArgs a{in, out, total, cdf, range, low, high, scale_in, norm, span, bl, dist_tukey, lam, loc, scale, rand,
                    philox_round_keys(seed), (uint32_t)offset, (uint32_t)(offset >> 32), index0, rand_out};
    SIMT_LAUNCH(blocks, 256, (hbr_map_kernel(a)));
    return 0;
}

}  // extern "C"
