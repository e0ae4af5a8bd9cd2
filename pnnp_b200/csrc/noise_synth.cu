// Fused physics-based noise synthesis for packed Bayer crops (sm_100a).
//
// Replaces the per-crop Python loops around generate_noisy_obs / generate_noisy_torch
// (data_process/process.py:591-673; callers data_process/syn_datasets.py:326-337 and
// trainer_SID.py:449-462): one launch for all crops, one HBM read of the clean tensor and one
// write of the noisy tensor (8 B per element), parameters from a small device table.
//
// Work decomposition: one warp per (crop, channel, row, 512-element segment).  Each lane owns
// up to four float4 groups (128-bit coalesced loads/stores, lane-interleaved).  The row-noise
// draw is keyed on the (crop, channel, row) index, so every lane of every warp that touches the
// row computes the same value — a broadcast by construction, no shuffle needed.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "abi_common.h"
#include "noise_kernels.cuh"

namespace pnnp {

static int ensure_poisson_table() {
    static bool done[64] = {};
    int dev = 0;
    PNNP_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail("noise_synth: device index out of range");
    if (done[dev]) return 0;
    static float host[kPoisTableFloats];
    build_poisson_table(host);
    PNNP_CUDA(cudaMemcpyToSymbol(g_pois_table, host, sizeof(host)));
    done[dev] = true;
    return 0;
}

static int check_common(const SynthArgs& a, int chain, bool replay) {
    if (!a.clean || !a.noisy || !a.table) return fail("noise_synth: null pointer");
    if (a.n <= 0 || a.c <= 0 || a.h <= 0 || a.w <= 0) return fail("noise_synth: empty shape");
    if (chain != PNNP_CHAIN_NUMPY && chain != PNNP_CHAIN_TORCH) return fail("noise_synth: unknown chain");
    if (chain == PNNP_CHAIN_TORCH) {
        // mirror the reference's failures (process.py:651,654,663)
        if (!(a.code & PNNP_CODE_P)) return fail("generate_noisy_torch: shot noise without 'p' is unsupported by the reference");
        if ((a.code & PNNP_CODE_G) && !(a.code & PNNP_CODE_B)) return fail("generate_noisy_torch: Tukey-lambda ('g') is NotImplemented in the reference");
        if (a.code & PNNP_CODE_D) return fail("generate_noisy_torch: 'd' is unsupported by the reference");
    }
    if ((a.code & PNNP_CODE_D) && a.c > 4) return fail("noise_synth: 'd' needs c <= 4");
    (void)replay;
    return 0;
}

template <bool DEBUG>
static int launch_synth(const SynthArgs& a, int chain, cudaStream_t st) {
    if (int e = check_common(a, chain, false)) return e;
    if (int e = ensure_poisson_table()) return e;
    int dev = 0, sms = 0;
    PNNP_CUDA(cudaGetDevice(&dev));
    PNNP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool vec = (a.w % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.clean) | reinterpret_cast<uintptr_t>(a.noisy)) % 16 == 0) &&
                     ((a.crop_id0 * (uint64_t)a.c * a.h * a.w) % 4 == 0);
    const int nseg = (a.w + kSeg - 1) / kSeg;
    const long long units = (long long)a.n * a.c * a.h * nseg;
    const long long want = (units + (kThreads / 32) - 1) / (kThreads / 32);
    const int blocks = (int)std::min<long long>(want, (long long)sms * 8);   // 8 CTAs of 256 threads per SM = 64 warps
    // Specialised path: the caller asserts (PNNP_CODE_UNIFORM_F64) that every table row has flags == K64|SIG64
    // (sample_params output); together with code == p|g|r|q, ori = clip = 0 and the vector layout this selects
    // the branch-free instantiation.  Anything else runs the generic kernel.
    const bool fast = vec && chain == PNNP_CHAIN_NUMPY && (a.code & PNNP_CODE_UNIFORM_F64) && (long long)a.n * a.c * a.h * a.w < (1ll << 31) &&
                      ((a.code & 0x3Fu) == (PNNP_CODE_P | PNNP_CODE_G | PNNP_CODE_R | PNNP_CODE_Q)) && !a.ori && !a.clip;
    SynthArgs b = a;
    b.code = a.code & 0x3Fu;
    b.rk = philox_round_keys(a.seed);
#define PNNP_LAUNCH(CH, V) noise_synth_kernel<CH, DEBUG, V><<<blocks, kThreads, 0, st>>>(b)
    if (fast) {
        const long long funits = (long long)a.n * a.c * a.h * ((a.w + kFastUnit - 1) / kFastUnit);
        const int exp_sw = DEBUG ? 0 : (getenv("PNNP_SYNTH_EXP") ? atoi(getenv("PNNP_SYNTH_EXP")) : 0);     // timing experiments only
        if (exp_sw) {
            const long long fwant = (funits + (kFastThreads / 32) - 1) / (kFastThreads / 32);
            const int fgrid = (int)std::min<long long>(fwant, (long long)sms * 3);
#define PNNP_EXP_CASE(E) case E: PNNP_CUDA(cudaFuncSetAttribute(noise_synth_fast_kernel<false, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmemBytes)); \
                                 noise_synth_fast_kernel<false, E><<<fgrid, kFastThreads, kFastSmemBytes, st>>>(b); break;
            switch (exp_sw) {
                PNNP_EXP_CASE(1) PNNP_EXP_CASE(2) PNNP_EXP_CASE(3) PNNP_EXP_CASE(5) PNNP_EXP_CASE(7) PNNP_EXP_CASE(8) PNNP_EXP_CASE(15)
                default: return fail("noise_synth: unknown PNNP_SYNTH_EXP combination");
            }
#undef PNNP_EXP_CASE
        } else {
            auto kern = noise_synth_fast_kernel<DEBUG>;
            const long long fwant = (funits + (kFastThreads / 32) - 1) / (kFastThreads / 32);
            static bool attr_done = false;
            if (!attr_done) {
                PNNP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmemBytes));
                attr_done = true;
            }
            kern<<<(int)std::min<long long>(fwant, (long long)sms * 3), kFastThreads, kFastSmemBytes, st>>>(b);
        }
    }
    else if (chain == PNNP_CHAIN_NUMPY) { if (vec) PNNP_LAUNCH(PNNP_CHAIN_NUMPY, 4); else PNNP_LAUNCH(PNNP_CHAIN_NUMPY, 1); }
    else                                { if (vec) PNNP_LAUNCH(PNNP_CHAIN_TORCH, 4); else PNNP_LAUNCH(PNNP_CHAIN_TORCH, 1); }
#undef PNNP_LAUNCH
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pnnp

using namespace pnnp;

extern "C" int pnnp_noise_synth(const float* clean, float* noisy, const pnnp_noise_params* table, int n, int c,
                                int h, int w, uint32_t code_bits, int chain, int ori, int clip, float post_lo,
                                float post_hi, uint64_t seed, uint64_t offset, uint64_t crop_id0, void* stream) {
    SynthArgs a{clean, noisy, table, n, c, h, w, code_bits, ori, clip, post_lo, post_hi, seed, offset, crop_id0, PhiloxKeys{},
                nullptr, nullptr, nullptr, nullptr};
    return launch_synth<false>(a, chain, (cudaStream_t)stream);
}

extern "C" int pnnp_noise_synth_debug(const float* clean, float* noisy, const pnnp_noise_params* table, int n,
                                      int c, int h, int w, uint32_t code_bits, int chain, int ori, int clip,
                                      float post_lo, float post_hi, uint64_t seed, uint64_t offset,
                                      uint64_t crop_id0, float* d_shot, float* d_read, float* d_rowz,
                                      double* d_q, void* stream) {
    SynthArgs a{clean, noisy, table, n, c, h, w, code_bits, ori, clip, post_lo, post_hi, seed, offset, crop_id0, PhiloxKeys{},
                d_shot, d_read, d_rowz, d_q};
    return launch_synth<true>(a, chain, (cudaStream_t)stream);
}

extern "C" int pnnp_noise_synth_replay(const float* clean, float* noisy, const pnnp_noise_params* table, int n,
                                       int c, int h, int w, uint32_t code_bits, int chain, int ori, int clip,
                                       float post_lo, float post_hi, const float* d_shot, const float* d_read,
                                       const float* d_rowz, const double* d_q, void* stream) {
    SynthArgs a{clean, noisy, table, n, c, h, w, code_bits, ori, clip, post_lo, post_hi, 0, 0, 0, PhiloxKeys{},
                const_cast<float*>(d_shot), const_cast<float*>(d_read), const_cast<float*>(d_rowz),
                const_cast<double*>(d_q)};
    if (int e = check_common(a, chain, true)) return e;
    const size_t total = (size_t)n * c * h * w;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    if (chain == PNNP_CHAIN_NUMPY) noise_replay_kernel<PNNP_CHAIN_NUMPY><<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    else noise_replay_kernel<PNNP_CHAIN_TORCH><<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
