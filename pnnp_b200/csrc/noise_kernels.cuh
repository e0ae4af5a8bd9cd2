// Kernels of the fused noise synthesis (launchers and the C ABI: noise_synth.cu).  They live in a header so that the CPU suite can
// compile this very source for the host and run whole CTAs of it — warp ballots, shuffles and shared memory included — on a
// lock-step fibre emulator (tests/emul/simt_host.h, tests/test_device_simt_on_cpu.py; test infrastructure only).
#pragma once
#include <algorithm>
#include <cmath>
#include "noise_core.cuh"

#ifndef PNNP_SMEM
#ifdef PNNP_HOST_EMUL
#define PNNP_SMEM static                 // one CTA at a time on the host: a static array is the CTA's shared memory
#else
#define PNNP_SMEM __shared__
#endif
#endif

namespace pnnp {

struct SynthArgs {
    const float* clean;
    float* noisy;
    const pnnp_noise_params* table;
    int n, c, h, w;
    uint32_t code;
    int ori, clip;
    float post_lo, post_hi;
    uint64_t seed, offset, crop_id0;
    PhiloxKeys rk;                   // round keys of `seed` (filled by launch_synth)
    // debug outputs / replay inputs (NULL when unused)
    float* d_shot; float* d_read; float* d_rowz; double* d_q;
};

// Poisson CDF table (noise_core.cuh: poisson_small_table), filled once per process by the host
__device__ float g_pois_table[kPoisTableFloats];
// the table's contents (host side; noise_synth.cu copies it to g_pois_table once per device)
inline void build_poisson_table(float* host) {
    for (int r = 0; r < kPoisRows; ++r) {
        const double lam = r / 16.0;
        double p = std::exp(-lam), F = p;
        float* row = host + r * kPoisStride;
        for (int z = 0; z < kPoisPad; ++z) row[z] = 0.f;
        for (int k = 0; k < kPoisCols; ++k) {
            if (k) { p *= lam / k; F += p; }
            row[kPoisPad + k] = (float)std::min(F, 1.0);
        }
    }
}
__device__ __forceinline__ void load_poisson_table(float* s_table) {
    for (int i = threadIdx.x; i < kPoisTableFloats; i += blockDim.x) s_table[i] = g_pois_table[i];
    __syncthreads();
}

#ifndef PNNP_HOST_EMUL
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
inline void prefetch_l2(const void*) {}
#endif

constexpr int kSeg = 512;          // elements per warp work unit
constexpr int kThreads = 256;

// Draws of one group of four elements (layout: noise_core.cuh).  shot[e] / mix[e] are words e of blocks sub 0 / 1; read[e] is
// the 32-bit word the read-noise sampler inverts (cell from mix, low bits from the refinement block only in the tails).
struct GroupDraws { uint32_t shot[4], mix[4], read[4]; };
__device__ __forceinline__ GroupDraws group_draws(const RngCtx& rng, uint64_t group) {
    GroupDraws g;
    const uint4 b0 = rng.block(group, kStreamElem, 0u), b1 = rng.block(group, kStreamElem, 1u);
    g.shot[0] = b0.x; g.shot[1] = b0.y; g.shot[2] = b0.z; g.shot[3] = b0.w;
    g.mix[0] = b1.x; g.mix[1] = b1.y; g.mix[2] = b1.z; g.mix[3] = b1.w;
    uint4 b2 = make_uint4(0u, 0u, 0u, 0u);
    if (read_cell_is_tail(b1.x >> 12) | read_cell_is_tail(b1.y >> 12) | read_cell_is_tail(b1.z >> 12) | read_cell_is_tail(b1.w >> 12))
        b2 = rng.block(group, kStreamElem, 2u);
    g.read[0] = read_word(b1.x, b2.x); g.read[1] = read_word(b1.y, b2.y);
    g.read[2] = read_word(b1.z, b2.z); g.read[3] = read_word(b1.w, b2.w);
    return g;
}

// Read-noise / quantisation draws only (the specialised kernel generates the shot words in an earlier phase).
__device__ __forceinline__ GroupDraws group_draws_read(const RngCtx& rng, uint64_t group) {
    GroupDraws g;
    const uint4 b1 = rng.block(group, kStreamElem, 1u);
    g.mix[0] = b1.x; g.mix[1] = b1.y; g.mix[2] = b1.z; g.mix[3] = b1.w;
    uint4 b2 = make_uint4(0u, 0u, 0u, 0u);
    if (read_cell_is_tail(b1.x >> 12) | read_cell_is_tail(b1.y >> 12) | read_cell_is_tail(b1.z >> 12) | read_cell_is_tail(b1.w >> 12))
        b2 = rng.block(group, kStreamElem, 2u);
    g.read[0] = read_word(b1.x, b2.x); g.read[1] = read_word(b1.y, b2.y);
    g.read[2] = read_word(b1.z, b2.z); g.read[3] = read_word(b1.w, b2.w);
    return g;
}

// One element given its draws (shot word, read word, mix word carrying the quantisation bits).  Returns the noisy value;
// optionally records the draws.
template <int CHAIN, bool DEBUG>
__device__ __forceinline__ float synth_one(float y, uint32_t w_shot, uint32_t w_read, uint32_t w_q, size_t lidx, int crop,
                                           int ch, const RowP& p, const SynthArgs& a, float rowz, float lam_tl,
                                           float inv_lam_tl, const float* pois_table) {
    const uint32_t code = a.code;
    float lam, d_shot;
    ScaleIn s;
    if (CHAIN == PNNP_CHAIN_NUMPY) { s = scale_in_numpy(y, p); lam = poisson_rate_numpy(s, p); }
    else { s.ysc32 = scale_in_torch(y, p); s.ysc64 = 0.0; lam = __fdiv_rn(s.ysc32, (float)p.K); }
    if (code & PNNP_CODE_P) d_shot = poisson_sample(lam, w_shot, pois_table);
    else d_shot = normal_icdf(w_shot);

    float d_read = 0.f;
    if (!(code & PNNP_CODE_B)) {
        if ((code & PNNP_CODE_G) && CHAIN == PNNP_CHAIN_NUMPY)
            d_read = tukey_lambda_ppf(w_read, lam_tl, inv_lam_tl) * (float)p.sigTL;
        else
            d_read = normal_icdf(w_read) * (float)p.sigGs;
    }
    float out;
    double dq = 0.0;
    if (CHAIN == PNNP_CHAIN_NUMPY) {
        // numpy: uniform(-0.5, 0.5) is float64; the 12-bit lattice value is exact in float64
        if (code & PNNP_CODE_Q) dq = quant_draw_f64(w_q);
        const double bias_c = (code & PNNP_CODE_D) ? a.table[crop].bias[ch & 3] : 0.0;
        out = tail_numpy(y, p, code, a.ori != 0, a.clip != 0, d_shot, d_read, rowz, dq, bias_c);
    } else {
        const float qu = quant_draw_f32(w_q);
        dq = (double)qu;
        out = tail_torch(p, code, a.ori != 0, a.clip != 0, d_shot, d_read, rowz, qu);
    }
    out = fminf(fmaxf(out, a.post_lo), a.post_hi);
    if (DEBUG) {
        if (a.d_shot) a.d_shot[lidx] = d_shot;
        if (a.d_read) a.d_read[lidx] = d_read;
        if (a.d_q) a.d_q[lidx] = dq;
    }
    return out;
}

template <int CHAIN, bool DEBUG, int VEC>
__global__ void __launch_bounds__(kThreads, 4) noise_synth_kernel(const SynthArgs a) {
    PNNP_SMEM float s_pois[kPoisTableFloats];
    load_poisson_table(s_pois);
    const int lane = threadIdx.x & 31;
    const long long warps_total = (long long)gridDim.x * (kThreads / 32);
    const long long warp_id = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int nseg = (a.w + kSeg - 1) / kSeg;
    const long long rows = (long long)a.n * a.c * a.h;
    const long long units = rows * nseg;
    const RngCtx rng{a.rk, (uint32_t)a.offset, (uint32_t)(a.offset >> 32)};
    const size_t crop_elems = (size_t)a.c * a.h * a.w;

    for (long long u = warp_id; u < units; u += warps_total) {
        const long long row = u / nseg;
        const int seg = (int)(u - row * nseg);
        const int crop = (int)(row / ((long long)a.c * a.h));
        const int ch = (int)((row / a.h) % a.c);
        const RowP p = load_row_params(a.table + crop);
        const float lam_tl = (float)p.lam;
        const float inv_lam_tl = lam_tl != 0.f ? 1.0f / lam_tl : 0.f;
        float rowz = 0.f;
        if (a.code & PNNP_CODE_R) {
            // keyed on the global (crop, channel, row) index: identical in every lane / warp that touches the row
            const uint64_t grow = a.crop_id0 * (uint64_t)a.c * a.h + (uint64_t)row;
            rowz = normal_icdf(rng.block(grow, kStreamRow, 0u).x);
            if (DEBUG && a.d_rowz && seg == 0 && lane == 0) a.d_rowz[row] = rowz;
        }
        const size_t row_base = (size_t)row * a.w;
        const uint64_t g_base = a.crop_id0 * (uint64_t)crop_elems + (uint64_t)row_base;
        const int x0 = seg * kSeg;
        if (VEC == 4) {
#pragma unroll 1
            for (int j = 0; j < kSeg / 128; ++j) {
                const int x = x0 + (j * 32 + lane) * 4;
                if (x >= a.w) break;
                const float4 y = __ldcs(reinterpret_cast<const float4*>(a.clean + row_base + x));
                const GroupDraws g = group_draws(rng, (g_base + x) >> 2);       // (g_base + x) % 4 == 0 on this path
                float4 o;
                o.x = synth_one<CHAIN, DEBUG>(y.x, g.shot[0], g.read[0], g.mix[0], row_base + x + 0, crop, ch, p, a, rowz, lam_tl, inv_lam_tl, s_pois);
                o.y = synth_one<CHAIN, DEBUG>(y.y, g.shot[1], g.read[1], g.mix[1], row_base + x + 1, crop, ch, p, a, rowz, lam_tl, inv_lam_tl, s_pois);
                o.z = synth_one<CHAIN, DEBUG>(y.z, g.shot[2], g.read[2], g.mix[2], row_base + x + 2, crop, ch, p, a, rowz, lam_tl, inv_lam_tl, s_pois);
                o.w = synth_one<CHAIN, DEBUG>(y.w, g.shot[3], g.read[3], g.mix[3], row_base + x + 3, crop, ch, p, a, rowz, lam_tl, inv_lam_tl, s_pois);
                __stcs(reinterpret_cast<float4*>(a.noisy + row_base + x), o);
            }
        } else {
#pragma unroll 1
            for (int x = x0 + lane; x < min(a.w, x0 + kSeg); x += 32) {
                const float y = a.clean[row_base + x];
                const uint64_t gi = g_base + x;
                const GroupDraws g = group_draws(rng, gi >> 2);
                const int e = (int)(gi & 3);
                uint32_t w0 = g.shot[0], w1 = g.read[0], w2 = g.mix[0];
                if (e == 1) { w0 = g.shot[1]; w1 = g.read[1]; w2 = g.mix[1]; }
                else if (e == 2) { w0 = g.shot[2]; w1 = g.read[2]; w2 = g.mix[2]; }
                else if (e == 3) { w0 = g.shot[3]; w1 = g.read[3]; w2 = g.mix[3]; }
                a.noisy[row_base + x] = synth_one<CHAIN, DEBUG>(y, w0, w1, w2, row_base + x, crop, ch, p, a, rowz, lam_tl, inv_lam_tl, s_pois);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Specialised kernel: NumPy chain, code = p|g|r|q, float64 K / sigR and python-float ratio (what sample_params returns),
// ori = clip = False, w % 4 == 0 — BASELINE configs[1].  Bit-identical to the generic kernel (same draws, same tail;
// test_specialised_kernel_is_bit_identical_to_replay_at_scale), organised around what the generic kernel wastes:
//
//  * The two Poisson samplers (exact CDF search below rate 10, Cornish-Fisher inversion above) are data-dependent branches
//    that a warp pays for one after the other whenever its 32 lanes disagree — and with per-pixel rates they always do;
//    inside the search every lane also waits for the slowest one.  Here a warp owns 512 consecutive elements of a row and
//    first *sorts them by sampler* through a 4 KB shared-memory queue (ballot + popc compaction: low-rate entries fill the
//    queue from the bottom, high-rate entries from the top), then runs each sampler over its part of the queue 32 entries at
//    a time with every lane active, and writes the count back in place.  Each lane keeps the queue positions of its 16
//    elements in registers and collects the counts afterwards.
//  * Per-crop constants (reciprocals for the Markstein divisions, float32 copies) are rebuilt only when the warp moves to
//    another crop; the row-noise draws of a warp's next 32 rows are generated in one go, one row per lane, and handed out
//    by shuffle (row noise is keyed on the global row index, so the value does not depend on who computes it).
//  * Phase 3 (read noise, quantisation, float64 tail, 128-bit streaming stores) needs only the counts, not the clean pixels.
// ------------------------------------------------------------------------------------------
constexpr int kFastUnit = 512;              // elements per warp work unit (16 per lane, four float4 groups)
constexpr int kFastThreads = 256;

constexpr int kFastQueueBytes = (kFastThreads / 32) * kFastUnit * 8, kFastPosBytes = (kFastThreads / 32) * kFastUnit * 2;
constexpr int kFastSmemBytes = kFastQueueBytes + kFastPosBytes + kPoisTableFloats * 4;

struct FastC {                              // per-crop constants
    float span32, ratio32, rratio32, invK32, sigTL32, lam_tl, inv_lam_tl;
    double K, span, rspan, lo, ratio, sigR;
};

// Three CTAs (24 warps) per SM at 79 registers without spills.  Measured alternatives (r01): four CTAs at 64 registers spill and
// run 15 % slower; reading the Poisson table through L1 instead of a per-CTA shared-memory copy (41 KB instead of 65 KB of
// shared memory) is 12 % slower at equal occupancy.
// EXP: timing experiments only (PNNP_SYNTH_EXP, never part of a result): bit 0 no Poisson samplers (count = round(rate)), bit 1 no
// Tukey-lambda quantile, bit 2 no sorting by sampler (queue slot = own slot), bit 3 no Philox rounds (counter mixed with two
// multiplies).  EXP = 0 is the product kernel; the others measure what each part costs ("measured floor").
template <int EXP>
__device__ __forceinline__ uint4 exp_block(const RngCtx& rng, uint64_t index, uint32_t stream, uint32_t sub) {
    if constexpr ((EXP & 8) != 0) {
        const uint32_t x = (uint32_t)index * 0x9E3779B9u + sub * 0x85EBCA6Bu + stream;
        return make_uint4(x, x * 0xC2B2AE35u, x ^ 0x27D4EB2Fu, x * 0x165667B1u);
    } else {
        return rng.block(index, stream, sub);
    }
}

template <bool DEBUG, int EXP = 0>
__global__ void __launch_bounds__(kFastThreads, 3) noise_synth_fast_kernel(const SynthArgs a) {
    // dynamic shared memory (kFastSmemBytes > 48 KB): [warps][512] uint2 queue | [warps][512] uint16 positions | Poisson table
    extern __shared__ __align__(16) uint8_t s_fast[];
    float* s_pois = reinterpret_cast<float*>(s_fast + kFastQueueBytes + kFastPosBytes);
    load_poisson_table(s_pois);
    const int lane = threadIdx.x & 31;
    uint2* q = reinterpret_cast<uint2*>(s_fast) + (threadIdx.x >> 5) * kFastUnit;
    uint16_t* qpos = reinterpret_cast<uint16_t*>(s_fast + kFastQueueBytes) + (threadIdx.x >> 5) * kFastUnit;   // queue position of every element
    const unsigned lt = (1u << lane) - 1u;
    // 32-bit index arithmetic (the launcher takes this kernel only below 2^31 elements): registers are what limits occupancy
    const int warps_total = (int)gridDim.x * (kFastThreads / 32);
    const int warp_id = (int)blockIdx.x * (kFastThreads / 32) + (int)(threadIdx.x >> 5);
    const int nseg = (a.w + kFastUnit - 1) / kFastUnit;
    const int rows_per_crop = a.c * a.h;
    const int units = a.n * rows_per_crop * nseg;
    const RngCtx rng{a.rk, (uint32_t)a.offset, (uint32_t)(a.offset >> 32)};

    FastC f = {};
    int cur_crop = -1;
    float rowz_batch = 0.f;
    int it = 0;
    for (int u = warp_id; u < units; u += warps_total, ++it) {
        if ((it & 31) == 0) {
            // row draws of this warp's next 32 units, one per lane
            const long long uu = (long long)u + (long long)lane * warps_total;
            if (uu < units) rowz_batch = normal_icdf(rng.block(a.crop_id0 * (uint64_t)rows_per_crop + (uint64_t)(uu / nseg), kStreamRow, 0u).x);
        }
        const float rowz = __shfl_sync(0xffffffffu, rowz_batch, it & 31);
        const int row = u / nseg;
        const int seg = u - row * nseg;
        const int crop = row / rows_per_crop;
        if (crop != cur_crop) {
            const pnnp_noise_params* t = a.table + crop;
            f.K = t->K; f.span = t->span; f.lo = t->clip_lo; f.ratio = t->ratio; f.sigR = t->sigR;
            f.rspan = __drcp_rn(f.span);
            f.span32 = (float)f.span; f.ratio32 = (float)f.ratio; f.rratio32 = __frcp_rn(f.ratio32);
            f.invK32 = (float)(1.0 / f.K); f.sigTL32 = (float)t->sigTL; f.lam_tl = (float)t->lam;
            f.inv_lam_tl = f.lam_tl != 0.f ? 1.0f / f.lam_tl : 0.f;
            cur_crop = crop;
        }
        if (DEBUG && a.d_rowz && seg == 0 && lane == 0) a.d_rowz[row] = rowz;
        const double row64 = __dmul_rn((double)rowz, f.sigR);
        const uint32_t row_base = (uint32_t)row * (uint32_t)a.w;
        const uint64_t g_base = a.crop_id0 * (uint64_t)rows_per_crop * (uint64_t)a.w + (uint64_t)row_base;
        const int x0 = seg * kFastUnit;

        // ---- phase 1: rates + shot words -> queue, sorted by sampler.  The j loops are deliberately NOT unrolled: the
        // kernel is latency-bound, not issue-bound, and a 4x unrolled body (4 Philox blocks per phase) overflows the
        // instruction cache once the warps of an SM spread over the three phases.
        int n_small = 0, n_large = 0;
        // all four 128-bit loads of the unit are issued before the first use (one exposed memory latency per unit instead of
        // four: the first multiply of a freshly loaded pixel was 13 % of all stall samples), and the next unit's lines are
        // requested into L2 while this unit computes
        float4 ybuf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + (j * 32 + lane) * 4;
            ybuf[j] = x < a.w ? __ldcs(reinterpret_cast<const float4*>(a.clean + row_base + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (u + warps_total < units) {
            const int un = u + warps_total, rown = un / nseg, xn0 = (un - rown * nseg) * kFastUnit;
            const float* nb = a.clean + (size_t)rown * a.w + xn0 + lane * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (xn0 + (j * 32 + lane) * 4 < a.w) prefetch_l2(nb + j * 128);
        }
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + (j * 32 + lane) * 4;
            const bool valid = x < a.w;
            const unsigned m_valid = __ballot_sync(0xffffffffu, valid);
            const float4 yv = j == 0 ? ybuf[0] : (j == 1 ? ybuf[1] : (j == 2 ? ybuf[2] : ybuf[3]));
            const uint4 b0 = exp_block<EXP>(rng, (g_base + x) >> 2, kStreamElem, 0u);      // (g_base + x) % 4 == 0 on this path
            const float ys[4] = {yv.x, yv.y, yv.z, yv.w};
            const uint32_t ws[4] = {b0.x, b0.y, b0.z, b0.w};
            uint32_t pq[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float ysc = div_rn_by_const(__fmul_rn(ys[e], f.span32), f.ratio32, f.rratio32);
                const float lam = ysc * f.invK32;
                if constexpr ((EXP & 4) != 0) {                               // experiment: no sorting, every entry in its own slot
                    pq[e] = (uint32_t)((j * 32 + lane) * 4 + e);
                    if (valid) q[pq[e]] = make_uint2(__float_as_uint(lam), ws[e]);
                    n_small = kFastUnit;
                } else {
                const bool small = lam < kPoissonSwitch;
                const unsigned m_small = __ballot_sync(0xffffffffu, small && valid);
                const unsigned m_large = m_valid & ~m_small;
                const int p_small = n_small + __popc(m_small & lt);
                const int p_large = kFastUnit - 1 - (n_large + __popc(m_large & lt));
                pq[e] = (uint32_t)(small ? p_small : p_large);
                if (valid) q[pq[e]] = make_uint2(__float_as_uint(lam), ws[e]);
                n_small += __popc(m_small);
                n_large += __popc(m_large);
                }
            }
            *reinterpret_cast<uint2*>(qpos + (j * 32 + lane) * 4) = make_uint2(pq[0] | (pq[1] << 16), pq[2] | (pq[3] << 16));
        }
        __syncwarp();
        // ---- phase 2: each sampler over its part of the queue, all lanes busy; the count replaces the rate in place
        for (int i = lane; i < n_small; i += 32) {
            const uint2 en = q[i];
            if constexpr ((EXP & 1) != 0) q[i].x = __float_as_uint(rintf(__uint_as_float(en.x)) + (float)(en.y >> 31));
            else if constexpr ((EXP & 4) != 0) q[i].x = __float_as_uint(poisson_sample(__uint_as_float(en.x), en.y, s_pois));
            else q[i].x = __float_as_uint(poisson_small_table(__uint_as_float(en.x), en.y, s_pois));
        }
        for (int i = lane; i < n_large; i += 32) {
            const uint2 en = q[kFastUnit - 1 - i];
            if constexpr ((EXP & 1) != 0) q[kFastUnit - 1 - i].x = __float_as_uint(rintf(__uint_as_float(en.x)) + (float)(en.y >> 31));
            else q[kFastUnit - 1 - i].x = __float_as_uint(poisson_large(__uint_as_float(en.x), en.y));
        }
        __syncwarp();
        // ---- phase 3: read noise + quantisation + tail (needs the counts, not the clean pixels)
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + (j * 32 + lane) * 4;
            if (x < a.w) {
                const uint64_t grp = (g_base + x) >> 2;
                const uint4 b1 = exp_block<EXP>(rng, grp, kStreamElem, 1u);
                const uint32_t mix[4] = {b1.x, b1.y, b1.z, b1.w};
                const uint2 pp = *reinterpret_cast<const uint2*>(qpos + (j * 32 + lane) * 4);
                const uint32_t pq[4] = {pp.x & 0xFFFFu, pp.x >> 16, pp.y & 0xFFFFu, pp.y >> 16};
                float d_read[4];
                if (read_cell_is_tail(b1.x >> 12) | read_cell_is_tail(b1.y >> 12) | read_cell_is_tail(b1.z >> 12) | read_cell_is_tail(b1.w >> 12)) {
                    // rare (2^-9 per group): some draw lies in the outer cells -> refinement block, general sampler
                    const uint4 b2 = rng.block(grp, kStreamElem, 2u);
                    const uint32_t rf[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) d_read[e] = tukey_lambda_ppf(read_word(mix[e], rf[e]), f.lam_tl, f.inv_lam_tl) * f.sigTL32;
                } else if constexpr ((EXP & 2) != 0) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) d_read[e] = (__uint_as_float(0x3F800000u | (mix[e] >> 9)) - 1.5f) * f.sigTL32;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) d_read[e] = tukey_lambda_ppf_body(mix[e], f.lam_tl, f.inv_lam_tl) * f.sigTL32;
                }
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float cnt = __uint_as_float(q[pq[e]].x);
                    const double dq = quant_draw_f64(mix[e]);
                    double A = __dmul_rn((double)cnt, f.K);
                    A = __dadd_rn(A, (double)d_read[e]);
                    A = __dadd_rn(A, row64);
                    A = __dadd_rn(A, dq);
                    const double z = clip_f64(div_rn_by_const(A, f.span, f.rspan), f.lo, 1.0);
                    o[e] = fminf(fmaxf((float)__dmul_rn(z, f.ratio), a.post_lo), a.post_hi);
                    if (DEBUG) {
                        const size_t lidx = (size_t)row_base + x + e;
                        if (a.d_shot) a.d_shot[lidx] = cnt;
                        if (a.d_read) a.d_read[lidx] = d_read[e];
                        if (a.d_q) a.d_q[lidx] = dq;
                    }
                }
                __stcs(reinterpret_cast<float4*>(a.noisy + row_base + x), make_float4(o[0], o[1], o[2], o[3]));
            }
        }
        __syncwarp();                       // the queue is reused by the next unit
    }
}

// Replay: same tails, draws read from memory.  One thread per element (test path, not tuned).
template <int CHAIN>
__global__ void noise_replay_kernel(const SynthArgs a) {
    const size_t total = (size_t)a.n * a.c * a.h * a.w;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / a.w;
        const int crop = (int)(row / ((size_t)a.c * a.h));
        const int ch = (int)((row / a.h) % a.c);
        const RowP p = load_row_params(a.table + crop);
        const float y = a.clean[i];
        const float d_shot = a.d_shot ? a.d_shot[i] : 0.f;
        const float d_read = a.d_read ? a.d_read[i] : 0.f;
        const float rowz = a.d_rowz ? a.d_rowz[row] : 0.f;
        const double dq = a.d_q ? a.d_q[i] : 0.0;
        float out;
        if (CHAIN == PNNP_CHAIN_NUMPY) {
            const double bias_c = (a.code & PNNP_CODE_D) ? a.table[crop].bias[ch & 3] : 0.0;
            out = tail_numpy(y, p, a.code, a.ori != 0, a.clip != 0, d_shot, d_read, rowz, dq, bias_c);
        } else {
            out = tail_torch(p, a.code, a.ori != 0, a.clip != 0, d_shot, d_read, rowz, (float)dq);
        }
        a.noisy[i] = fminf(fmaxf(out, a.post_lo), a.post_hi);
    }
}

}  // namespace pnnp
