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

// Poisson CDF table (noise_core.cuh: poisson_small_table_k), filled once per process by the host
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
//    that a warp pays for one after the other whenever its 32 lanes disagree — and with per-pixel rates they always do.
//    Here a warp owns 512 consecutive elements of a row and first *sorts them by sampler* through a 4 KB shared-memory queue
//    (low-rate entries fill it from the bottom, high-rate entries from the top), then runs each sampler over its part of the
//    queue 32 entries at a time with every lane active.  An entry is (rate, shot word); the 9 low bits of the word, which no
//    sampler reads (noise_core.cuh: shot_word), carry the element's position in the unit, so the count goes straight to
//    the element's slot of a 2 KB count array and phase 3 fetches its four counts with one conflict-free 128-bit load.
//  * Queue positions come from ONE warp scan per unit: a lane counts the low-rate elements among its 16 (a bit mask), the
//    packed (low, high) counts go through five shuffle steps, and the lane's entries then take consecutive slots.  (Round 1
//    used a ballot and four POPC per element; POPC, the conversions and MUFU share the XU pipe — four lanes per SM
//    sub-partition and clock — which that kernel kept 38 % busy.)
//  * Per-crop constants (reciprocals for the Markstein divisions, float32 copies, the folded clip bounds) are rebuilt only
//    when the warp moves to another crop; row / crop indices advance incrementally (no integer division per unit); the
//    row-noise draws of a warp's next 32 rows are generated in one go, one row per lane, and handed out by shuffle (row
//    noise is keyed on the global row index, so the value does not depend on who computes it).
//  * Phase 3 (read noise, quantisation, float64 tail, 128-bit streaming stores) needs only the counts, not the clean pixels.
// ------------------------------------------------------------------------------------------
constexpr int kFastUnit = 512;              // elements per warp work unit (16 per lane, four float4 groups)
constexpr int kFastThreads = 256;
static_assert(kFastUnit - 1 <= (int)kShotPosMask, "an element's position in the unit travels in the unused low bits of its shot word");

constexpr int fast_queue_bytes(int threads) { return (threads / 32) * kFastUnit * 8; }
constexpr int fast_count_bytes(int threads) { return (threads / 32) * kFastUnit * 4; }
constexpr int fast_smem_bytes(int threads) { return fast_queue_bytes(threads) + fast_count_bytes(threads) + kPoisTableFloats * 4; }
constexpr int kFastSmemBytes = fast_smem_bytes(kFastThreads);

// Three CTAs (24 warps) per SM.  Measured alternatives: four CTAs at 64 registers spill and run 15 % slower (r01); two CTAs of 448
// threads at 72 registers (28 warps) 6 % slower (r02: 575 against 544 us); reading the Poisson table through L1 instead of a
// per-CTA shared-memory copy is 12 % slower at equal occupancy (r01).
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

// Queue stores of phase 1.  `sp` / `lp` are shared-memory byte addresses; an entry goes to the low-rate end (and moves `sp` up)
// or to the high-rate end (and moves `lp` down): one predicate, one select, two predicated adds and the store.
#ifndef PNNP_HOST_EMUL
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void queue_put_at(uint32_t addr, float lam, uint32_t word) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(__float_as_uint(lam)), "r"(word) : "memory");
}
__device__ __forceinline__ void queue_put(uint32_t& sp, uint32_t& lp, float lam, uint32_t word) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 a;\n\t"
        "setp.lt.f32 p, %2, 0f41200000;\n\t"          // lam < 10
        "selp.b32 a, %0, %1, p;\n\t"
        "st.shared.v2.b32 [a], {%3, %4};\n\t"
        "@p add.u32 %0, %0, 8;\n\t"
        "@!p sub.u32 %1, %1, 8;\n\t}"
        : "+r"(sp), "+r"(lp) : "f"(lam), "r"(__float_as_uint(lam)), "r"(word) : "memory");
}
#else
extern uint8_t s_fast[];              // the emulated CTA's dynamic shared memory (tests/emul/simt_kernels_host.cpp)
inline uint32_t smem_addr(const void* p) { return (uint32_t)(uintptr_t)((const uint8_t*)p - s_fast); }
inline void queue_put_at(uint32_t addr, float lam, uint32_t word) { *reinterpret_cast<uint2*>(s_fast + addr) = make_uint2(__float_as_uint(lam), word); }
inline void queue_put(uint32_t& sp, uint32_t& lp, float lam, uint32_t word) {
    const bool small = lam < kPoissonSwitch;
    queue_put_at(small ? sp : lp, lam, word);
    if (small) sp += 8; else lp -= 8;
}
#endif

template <bool DEBUG, int EXP = 0, int THREADS = kFastThreads, int CTAS = 3>
__global__ void __launch_bounds__(THREADS, CTAS) noise_synth_fast_kernel(const SynthArgs a) {
    // dynamic shared memory (fast_smem_bytes(THREADS) > 48 KB): [warps][512] uint2 queue | [warps][512] int counts | Poisson table
    extern __shared__ __align__(16) uint8_t s_fast[];
    constexpr int kFastWarps = THREADS / 32, kFastQueueBytes = fast_queue_bytes(THREADS), kFastCountBytes = fast_count_bytes(THREADS);
    float* s_pois = reinterpret_cast<float*>(s_fast + kFastQueueBytes + kFastCountBytes);
    load_poisson_table(s_pois);
    const int lane = threadIdx.x & 31;
    uint2* q = reinterpret_cast<uint2*>(s_fast) + (threadIdx.x >> 5) * kFastUnit;
    int* cnt_s = reinterpret_cast<int*>(s_fast + kFastQueueBytes) + (threadIdx.x >> 5) * kFastUnit;   // counts, element order
    // 32-bit index arithmetic (the launcher takes this kernel only below 2^31 elements): registers are what limits occupancy
    const int warps_total = (int)gridDim.x * kFastWarps;
    const int warp_id = (int)blockIdx.x * kFastWarps + (int)(threadIdx.x >> 5);
    const int nseg = (a.w + kFastUnit - 1) / kFastUnit;
    const int rows_per_crop = a.c * a.h;
    const int units = a.n * rows_per_crop * nseg;
    const RngCtx rng{a.rk, (uint32_t)a.offset, (uint32_t)(a.offset >> 32)};
    // Philox group index (global element index / 4) of the tensor's first element; opaque to the compiler, which otherwise
    // re-derives the 64-bit product in every trip of the loops below instead of keeping two registers
    uint64_t group0 = (a.crop_id0 * (uint64_t)rows_per_crop * (uint64_t)a.w) >> 2;          // the product is a multiple of 4 on this path
#ifndef PNNP_HOST_EMUL
    asm volatile("" : "+l"(group0));
#endif
    // A warp owns a CONTIGUOUS range of units (rows follow each other, so the per-crop constants below are rebuilt once or
    // twice per warp — with a grid-stride walk and more warps than rows per crop EVERY unit was in another crop: ~100
    // instructions and a float64 reciprocal per unit); (row, segment, crop, row inside the crop) advance by increment and carry.
    const int u_begin = (int)((long long)units * warp_id / warps_total), u_end = (int)((long long)units * (warp_id + 1) / warps_total);
    int row = u_begin / nseg, seg = u_begin - row * nseg;
    int crop = row / rows_per_crop, crop_row = row - crop * rows_per_crop;

    FastC f = {};
    int cur_crop = -1;
    float rowz_batch = 0.f;
    int it = 0;
    for (int u = u_begin; u < u_end; ++u, ++it) {
        if ((it & 31) == 0) {
            // row draws of this warp's next 32 units, one per lane
            const int uu = u + lane;
            if (uu < u_end) rowz_batch = normal_icdf(rng.block(a.crop_id0 * (uint64_t)rows_per_crop + (uint64_t)(uu / nseg), kStreamRow, 0u).x);
        }
        const float rowz = __shfl_sync(0xffffffffu, rowz_batch, it & 31);
        if (crop != cur_crop) {
            f = fast_constants(a.table + crop, a.post_lo, a.post_hi);
            cur_crop = crop;
        }
        if (DEBUG && a.d_rowz && seg == 0 && lane == 0) a.d_rowz[row] = rowz;
        const double row64 = __dmul_rn((double)rowz, f.sigR);
        const uint32_t row_base = (uint32_t)row * (uint32_t)a.w;
        const int x0 = seg * kFastUnit;
        const uint64_t group_u = group0 + (uint64_t)((row_base + (uint32_t)x0) >> 2) + (uint64_t)lane;   // group of this lane's first float4

        // ---- phase 1: rates, sampler of every element, ONE scan for the queue positions, shot words -> queue.
        // All four 128-bit loads of the unit are issued before the first use (one exposed memory latency per unit), and the
        // next unit's lines are requested into L2 while this unit computes.
        float4 ybuf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + (j * 32 + lane) * 4;
            ybuf[j] = x < a.w ? __ldcs(reinterpret_cast<const float4*>(a.clean + row_base + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (u + 1 < u_end) {
            int rown = row, segn = seg + 1;
            if (segn >= nseg) { segn = 0; ++rown; }
            const float* nb = a.clean + (size_t)rown * a.w + segn * kFastUnit + lane * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (segn * kFastUnit + (j * 32 + lane) * 4 < a.w) prefetch_l2(nb + j * 128);
        }
        // rates: counted by sampler here and parked in the (still unused) count array, so that no 16 values stay in registers
        // across the scan and the Philox blocks below
        int ns_own = 0, n_own = 0;            // low-rate / valid elements of this lane
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool valid = x0 + (j * 32 + lane) * 4 < a.w;
            n_own += valid ? 4 : 0;
            float4 l;
            l.x = fast_rate(f, ybuf[j].x); l.y = fast_rate(f, ybuf[j].y); l.z = fast_rate(f, ybuf[j].z); l.w = fast_rate(f, ybuf[j].w);
            ns_own += (valid && l.x < kPoissonSwitch) ? 1 : 0;
            ns_own += (valid && l.y < kPoissonSwitch) ? 1 : 0;
            ns_own += (valid && l.z < kPoissonSwitch) ? 1 : 0;
            ns_own += (valid && l.w < kPoissonSwitch) ? 1 : 0;
            *reinterpret_cast<float4*>(cnt_s + (j * 32 + lane) * 4) = l;
        }
        int n_small, n_large;
        uint32_t sp, lp;                      // next low-rate slot (upwards) / high-rate slot (downwards) of this lane, as shared-memory byte addresses
        const uint32_t q_addr = smem_addr(q);
        if constexpr ((EXP & 4) != 0) {       // experiment: no sorting, every entry in its own slot, one sampler loop with both samplers
            n_small = kFastUnit; n_large = 0; sp = q_addr; lp = q_addr;
        } else {
            const uint32_t own = (uint32_t)ns_own | ((uint32_t)(n_own - ns_own) << 16);
            uint32_t incl = own;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                incl += lane >= d ? up : 0u;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            n_small = (int)(tot & 0xFFFFu); n_large = (int)(tot >> 16);
            sp = q_addr + 8u * ((incl - own) & 0xFFFFu);
            lp = q_addr + 8u * (uint32_t)(kFastUnit - 1 - (int)((incl - own) >> 16));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + (j * 32 + lane) * 4;
            if (x < a.w) {
                const uint4 b0 = exp_block<EXP>(rng, group_u + (uint64_t)(j * 32), kStreamElem, 0u);
                const uint32_t ws[4] = {b0.x, b0.y, b0.z, b0.w};
                const float4 l4 = *reinterpret_cast<const float4*>(cnt_s + (j * 32 + lane) * 4);      // own store: no barrier needed
                const float lam[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t pos_in_unit = (uint32_t)(lane * 4) | (uint32_t)(j * 128 + e);
                    const uint32_t word = (ws[e] & ~kShotPosMask) | pos_in_unit;
                    if constexpr ((EXP & 4) != 0) queue_put_at(q_addr + 8u * pos_in_unit, lam[e], word);
                    else queue_put(sp, lp, lam[e], word);
                }
            }
        }
        __syncwarp();
        // ---- phase 2: each sampler over its part of the queue, all lanes busy; the count goes to the element's slot
        // Two entries per lane and trip: the samplers are chains of dependent shared-memory loads / special-function results and
        // the kernel runs 6 warps per scheduler, so a second independent chain fills the slots the first one waits in (an odd
        // last trip repeats its entry: same count to the same slot).
        if constexpr ((EXP & 5) != 0) {
            for (int i = lane; i < n_small; i += 32) {
                const uint2 en = q[i];
                const float l = __uint_as_float(en.x);
                int k;
                if constexpr ((EXP & 1) != 0) k = (int)(rintf(l) + (float)(en.y >> 31));
                else k = (int)poisson_sample(l, en.y, s_pois);
                cnt_s[en.y & kShotPosMask] = k;
            }
            for (int i = lane; i < n_large; i += 32) {
                const uint2 en = q[kFastUnit - 1 - i];
                cnt_s[en.y & kShotPosMask] = (int)(rintf(__uint_as_float(en.x)) + (float)(en.y >> 31));
            }
        } else {
            for (int i = lane; i < n_small; i += 64) {
                const uint2 e0 = q[i], e1 = q[i + 32 < n_small ? i + 32 : i];
                PoisSmall s0 = poisson_small_search(__uint_as_float(e0.x), e0.y, s_pois);
                PoisSmall s1 = poisson_small_search(__uint_as_float(e1.x), e1.y, s_pois);
                const int k0 = poisson_small_finish(s0, __uint_as_float(e0.x), e0.y);
                const int k1 = poisson_small_finish(s1, __uint_as_float(e1.x), e1.y);
                cnt_s[e0.y & kShotPosMask] = k0;
                cnt_s[e1.y & kShotPosMask] = k1;
            }
            for (int i = lane; i < n_large; i += 64) {
                const uint2 e0 = q[kFastUnit - 1 - i], e1 = q[kFastUnit - 1 - (i + 32 < n_large ? i + 32 : i)];
                const int k0 = poisson_large_k(__uint_as_float(e0.x), shot_word(e0.y));
                const int k1 = poisson_large_k(__uint_as_float(e1.x), shot_word(e1.y));
                cnt_s[e0.y & kShotPosMask] = k0;
                cnt_s[e1.y & kShotPosMask] = k1;
            }
        }
        __syncwarp();
        // ---- phase 3: read noise + quantisation + tail (needs the counts, not the clean pixels).  Deliberately NOT unrolled:
        // a 4x unrolled body (4 Philox blocks + 16 tails) overflows the instruction cache once the warps of an SM spread over
        // the three phases.
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + (j * 32 + lane) * 4;
            if (x < a.w) {
                const uint64_t grp = group_u + (uint64_t)(j * 32);
                const uint4 b1 = exp_block<EXP>(rng, grp, kStreamElem, 1u);
                const uint32_t mix[4] = {b1.x, b1.y, b1.z, b1.w};
                const uint4 c4 = *reinterpret_cast<const uint4*>(cnt_s + (j * 32 + lane) * 4);
                const int cnt[4] = {(int)c4.x, (int)c4.y, (int)c4.z, (int)c4.w};
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
                    for (int e = 0; e < 4; ++e) d_read[e] = tukey_lambda_ppf_body_pow(mix[e], f.lam_tl, f.inv_lam_tl) * f.sigTL32;
                }
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double dq = quant_draw_f64(mix[e]);
                    o[e] = fast_tail(f, cnt[e], d_read[e], row64, dq);
                    if (DEBUG) {
                        const size_t lidx = (size_t)row_base + x + e;
                        if (a.d_shot) a.d_shot[lidx] = (float)cnt[e];
                        if (a.d_read) a.d_read[lidx] = d_read[e];
                        if (a.d_q) a.d_q[lidx] = dq;
                    }
                }
                __stcs(reinterpret_cast<float4*>(a.noisy + row_base + x), make_float4(o[0], o[1], o[2], o[3]));
            }
        }
        __syncwarp();                       // the queue and the count array are reused by the next unit
        // next unit of this warp
        if (++seg >= nseg) { seg = 0; ++row; if (++crop_row >= rows_per_crop) { crop_row = 0; ++crop; } }
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
