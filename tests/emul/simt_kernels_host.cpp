// Whole CTAs of the warp-collective kernels compiled for the HOST and run on the lock-step fibre emulator (simt_host.h) — TEST
// INFRASTRUCTURE ONLY (tests/test_device_simt_on_cpu.py).  Same kernel source as the GPU build (csrc/noise_kernels.cuh).
#define PNNP_HOST_EMUL 1
#include "cuda_host_shim.h"
#include "simt_host.h"
#include "../../pnnp_b200/csrc/noise_kernels.cuh"

namespace pnnp { alignas(16) uint8_t s_fast[kFastSmemBytes]; }     // the fast kernel's dynamic shared memory

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

}  // extern "C"
