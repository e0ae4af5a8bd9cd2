// First layer of both networks fused with the pack boundary: packed Bayer planes NCHW fp32 (<= 4 channels, what the noise
// synthesis / raw2bayer write: data_process/process.py:625-631, utils/isp_ops.py:84-96) -> conv3x3 pad 1 + bias + LeakyReLU / ReLU
// -> NHWC bf16 (archs/Unet.py:55 conv1_1, archs/ResUnet.py conv_in).  Replaces two launches of the general path — the
// NCHW fp32 -> NHWC16 bf16 conversion (145 MB of traffic per Sony frame) and a conv that reads 16-channel-padded pixels (4x the
// real input bytes) — by ONE kernel that reads the 4 real channels once.
//
// GEMM view per tile: M = 128 output pixels (8 x 16), N = cout (16..64), K = 9 taps x 4 channels = 36, zero-padded to 48 = three
// tcgen05.mma K16 steps.  A 4-channel pixel is 8 bytes of bf16 — below TMA's 16-byte inner-box minimum and not a legal core-matrix
// row — so the A operand is built by the CTA's threads (im2col in shared memory): the 10 x 18 x 4 fp32 halo tile is loaded with
// coalesced, bounds-predicated loads (= the conv's zero padding), each thread converts the 36 values of its pixel to bf16 and
// writes its row of the canonical K-major SWIZZLE_32B layout (the layout the general kernel's kc = 16 stages use), then one
// thread issues the three MMAs; the accumulator (cout fp32 columns of TMEM) is drained with tcgen05.ld by the same four warps.
// The next tile's global loads are in flight while the current tile is built, multiplied and stored; several CTAs per SM overlap
// their phases.  HBM traffic = 16 B read + 2 * cout B written per pixel: HBM-bound.
#ifndef PNNP_HOST_EMUL
#include <cuda.h>
#include <cuda_bf16.h>
#endif
#include <algorithm>
#include "abi_common.h"
#include "tc_common.cuh"
#include "../../include/pnnp_b200.h"

namespace pnnp {

constexpr int kFirstThreads = 128;
constexpr int kFirstTileH = 8, kFirstTileW = 16;
constexpr int kFirstHaloH = kFirstTileH + 2, kFirstHaloW = kFirstTileW + 2;          // 10 x 18
constexpr int kFirstPlane = kFirstHaloH * kFirstHaloW;                               // 180 floats per channel
constexpr int kFirstSlots = 4 * kFirstPlane;                                         // 720 halo-tile elements, 6 per thread (48 idle)
constexpr int kFirstABlock = 128 * 32;                                               // one K16 step of A: 128 rows x 32 bytes
constexpr int kFirstBBlock = 64 * 32;                                                // one K16 step of B: up to 64 rows x 32 bytes
constexpr int kFirstSmem = 1024 /*align slack*/ + 3 * kFirstABlock + 3 * kFirstBBlock + 768 * 4 + 64 * 4 + 64;

struct FirstParams {
    const float* in; const float* w; const float* bias; __nv_bfloat16* out;
    int n, cin, H, W, cout, tiles_x, tiles_y;
    float slope;                 // activation as max(v, v * slope): LeakyReLU 0.2 / ReLU 0 / identity 1
    int dbg;                     // timing experiments only (PNNP_FIRST_DBG): 1 no MMA round trip, 2 no global loads, 4 no stores, 8 no im2col
    int* err;
};

// byte offset of (row r, 16-byte chunk c) inside a K-major SWIZZLE_32B block whose base is 256-byte aligned
__device__ __forceinline__ uint32_t swz32(uint32_t r, uint32_t c) { return r * 32u + ((c ^ ((r >> 2) & 1u)) << 4); }

__device__ __forceinline__ uint16_t bf1(float a) {
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    return *reinterpret_cast<const uint16_t*>(&h);
}
__device__ __forceinline__ uint32_t bf2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(kFirstThreads, 8) conv_first_kernel(const FirstParams p) {
    // dynamic shared memory, carved: [A: 3 x 4 KB][B: 3 x 2 KB][halo tile: 768 floats][bias: 64 floats][mbarrier][TMEM slot]
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by POINTER arithmetic on the __shared__ array (not through an integer cast): the compiler keeps the
    // shared address space of everything carved from it, so plain loads / stores of these pointers are LDS / STS instead of
    // generic LD / ST (which go through the global-memory instruction queue: stall reason lg_throttle in the r02 capture)
    uint8_t* s_ab = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_in = reinterpret_cast<float*>(s_ab + 3 * kFirstABlock + 3 * kFirstBBlock);
    float* s_bias = s_in + 768;
    uint64_t& s_bar = *reinterpret_cast<uint64_t*>(s_bias + 64);
    uint32_t& s_tmem = *reinterpret_cast<uint32_t*>(s_bias + 64 + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sA = s_ab;
    uint8_t* sB = s_ab + 3 * kFirstABlock;
    const int tmem_cols = p.cout <= 32 ? 32 : 64;

    // ---- once per CTA: weights -> bf16 B operand (row = output channel, k = tap * 4 + channel), bias, barrier, TMEM
    for (int i = tid; i < 3 * kFirstBBlock / 16; i += kFirstThreads) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < kFirstABlock / 16; i += kFirstThreads) reinterpret_cast<uint4*>(sA + 2 * kFirstABlock)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int i = tid; i < p.cout * 36; i += kFirstThreads) {
        const int nrow = i / 36, k = i - nrow * 36, c = k & 3, tap = k >> 2;
        if (c < p.cin) {
            const __nv_bfloat16 v = __float2bfloat16_rn(p.w[(nrow * p.cin + c) * 9 + tap]);
            const int kb = k >> 4, kk = k & 15;
            *reinterpret_cast<__nv_bfloat16*>(sB + kb * kFirstBBlock + swz32((uint32_t)nrow, (uint32_t)(kk >> 3)) + (kk & 7) * 2) = v;
        }
    }
    for (int i = tid; i < 64; i += kFirstThreads) s_bias[i] = (i < p.cout && p.bias) ? p.bias[i] : 0.f;
    if (tid == 0) { mbar_init(smem_u32(&s_bar), 1); fence_mbarrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), (uint32_t)tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dhi = umma_desc_hi(32);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB), bar = smem_u32(&s_bar);

    // everything the loop needs in registers: the asm statements carry "memory" clobbers, so p.field would be re-read per tile
    const int H = p.H, W = p.W, cin = p.cin, cout = p.cout, tiles_x = p.tiles_x, tiles_y = p.tiles_y;
    const float slope = p.slope;
    const int dbg = p.dbg;
    const float* const in = p.in;
    __nv_bfloat16* const out = p.out;
    int* const err = p.err;
    const int total = p.n * tiles_y * tiles_x;
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;           // contiguous tile range: neighbouring halos meet in L2
    const int t_begin = min(total, (int)blockIdx.x * per), t_end = min(total, t_begin + per);
    const size_t plane = (size_t)H * W;

    // Halo-tile loads: warp w fetches channel w's 10 x 18 window, lane l the elements idx = l + 32 j (j < 6, idx < 180) — row
    // idx / 18, column idx % 18, fixed per thread — so a tile costs six loads from one base pointer plus a per-thread offset.
    const int lane = tid & 31;
    int e_off[6], e_row[6], e_col[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const int idx = lane + 32 * j;
        e_row[j] = idx / kFirstHaloW;
        e_col[j] = idx - e_row[j] * kFirstHaloW;
        e_off[j] = e_row[j] * W + e_col[j];
    }
    const bool ch_on = warp < cin;
    const bool last_on = lane + 160 < kFirstPlane;                           // slot j = 5 exists for lanes 0..19
    float pre[6];
    // (img, ty, tx) of the tile being prefetched, advanced by increment-and-carry
    int n_img = t_begin / (tiles_y * tiles_x), n_ty, n_tx;
    { const int r = t_begin - n_img * (tiles_y * tiles_x); n_ty = r / tiles_x; n_tx = r - n_ty * tiles_x; }
    auto prefetch = [&]() {
        const int y0 = n_ty * kFirstTileH - 1, x0 = n_tx * kFirstTileW - 1;
        const float* base = in + ((size_t)n_img * cin + warp) * plane + (long long)y0 * W + x0;
        if (!ch_on || (dbg & 2)) {
#pragma unroll
            for (int j = 0; j < 6; ++j) pre[j] = 0.f;
        } else if (y0 >= 0 && x0 >= 0 && y0 + kFirstHaloH <= H && x0 + kFirstHaloW <= W) {      // interior tile: no bounds tests
#pragma unroll
            for (int j = 0; j < 5; ++j) pre[j] = __ldg(base + e_off[j]);
            pre[5] = last_on ? __ldg(base + e_off[5]) : 0.f;
        } else {                                                                                   // border: zero padding by predicate
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int gy = y0 + e_row[j], gx = x0 + e_col[j];
                const bool ok = (j < 5 || last_on) && gy >= 0 && gy < H && gx >= 0 && gx < W;
                pre[j] = ok ? __ldg(base + e_off[j]) : 0.f;
            }
        }
        if (++n_tx == tiles_x) { n_tx = 0; if (++n_ty == tiles_y) { n_ty = 0; ++n_img; } }
    };
    int c_img = n_img, c_ty = n_ty, c_tx = n_tx;                             // the tile being computed
    if (t_begin < t_end) prefetch();
    const int my = tid >> 4, mx = tid & 15;                                  // this thread's pixel inside the tile = accumulator row
    float* const s_mine = s_in + warp * kFirstPlane + lane;
    const float* const q0 = s_in + my * kFirstHaloW + mx;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint64_t slope2 = f2_pack(slope, slope), zero2 = f2_pack(0.0f, 0.0f);
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
#pragma unroll
        for (int j = 0; j < 5; ++j) s_mine[32 * j] = pre[j];
        if (last_on) s_mine[160] = pre[5];
        __syncthreads();
        if (t + 1 < t_end) prefetch();                                       // in flight while this tile is built, multiplied, stored
        // ---- im2col: this pixel's 9 taps x 4 channels -> bf16 -> its row of the three K16 blocks
        uint32_t pk[9][2];
        if (!(dbg & 8)) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* q = q0 + ky * kFirstHaloW + kx;
                pk[ky * 3 + kx][0] = bf2(q[0], q[kFirstPlane]);
                pk[ky * 3 + kx][1] = bf2(q[2 * kFirstPlane], q[3 * kFirstPlane]);
            }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)                                       // 16-byte chunk ch of the row = taps 2 ch, 2 ch + 1
            *reinterpret_cast<uint4*>(sA + (ch >> 1) * kFirstABlock + swz32((uint32_t)tid, (uint32_t)(ch & 1))) =
                make_uint4(pk[2 * ch][0], pk[2 * ch][1], pk[2 * ch + 1][0], pk[2 * ch + 1][1]);
        *reinterpret_cast<uint4*>(sA + 2 * kFirstABlock + swz32((uint32_t)tid, 0u)) = make_uint4(pk[8][0], pk[8][1], 0u, 0u);
        }
        if (!(dbg & 1)) fence_proxy_async();                                 // generic-proxy writes -> visible to the tensor core
        tc_fence_before();
        __syncthreads();
        if (tid == 0 && !(dbg & 1)) {
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < 3; ++kb)
                tc_mma_bf16(tmem_base, umma_desc(dhi, a0 + kb * kFirstABlock), umma_desc(dhi, b0 + kb * kFirstBBlock), idesc, kb ? 1u : 0u);
            tc_commit(bar);
        }
        // ---- epilogue: bias + activation -> bf16 NHWC
        const int y = c_ty * kFirstTileH + my, x = c_tx * kFirstTileW + mx;
        const bool valid = y < H && x < W;
        __nv_bfloat16* op = out + (((size_t)c_img * H + y) * (size_t)W + x) * cout;
        if (++c_tx == tiles_x) { c_tx = 0; if (++c_ty == tiles_y) { c_ty = 0; ++c_img; } }
        if (!(dbg & 1)) { mbar_wait(bar, phase, err, 301); phase ^= 1; }
        tc_fence_after();
        for (int j = 0; j < cout / 16; ++j) {
            uint32_t v[16];
            if (!(dbg & 1)) { tc_ld16(taddr + j * 16, v); tc_ld_wait(); }
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(pre[i % 6]);
            }
            uint32_t o[8];
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + j * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {                                    // packed pairs: the same IEEE add / fma, half the instructions
                const float4 bb = b4[i];
                const uint64_t a01 = f2_add(f2_pack(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), f2_pack(bb.x, bb.y));
                const uint64_t a23 = f2_add(f2_pack(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), f2_pack(bb.z, bb.w));
                float a0_, a1_, a2_, a3_, m0, m1, m2, m3;
                f2_unpack(a01, a0_, a1_); f2_unpack(a23, a2_, a3_);
                f2_unpack(f2_fma(a01, slope2, zero2), m0, m1); f2_unpack(f2_fma(a23, slope2, zero2), m2, m3);
                o[2 * i] = bf2(fmaxf(a0_, m0), fmaxf(a1_, m1));
                o[2 * i + 1] = bf2(fmaxf(a2_, m2), fmaxf(a3_, m3));
            }
            if (valid && !(dbg & 4)) st_global_256(op + j * 16, o);
        }
        tc_fence_before();                                                   // the next MMA overwrites the accumulator: ordered by the barrier
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)tmem_cols); }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Warp-specialised form (the default).  The single-role kernel above runs load -> im2col -> MMA -> drain -> store as ONE dependency
// chain per CTA; its r02 timing experiments (PNNP_FIRST_DBG, 64 crops: 558 us; no stores 387; no im2col 387; no loads 490; no MMA
// round trip 490; all four off 148) show the parts simply ADD.  Here they overlap inside a CTA:
//   warps 0-3  producers: halo-tile loads (prefetched one tile ahead) -> shared fp32 tile (double-buffered) -> im2col rows of A
//              (double-buffered) -> fence.proxy.async -> arrive a_full[b];
//   warp  8    MMA issuer: waits a_full[b] and t_empty[b], three tcgen05.mma into accumulator b, commits to a_empty[b] and t_full[b];
//   warps 4-7  epilogue: wait t_full[b] -> tcgen05.ld -> release the accumulator -> bias + activation -> 64-byte stores.
// Same arithmetic, same operand layouts, same results bit for bit as the single-role kernel (tested).
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int kWsThreads = 288;
constexpr int kWsSmem = 1024 + 2 * 3 * kFirstABlock + 3 * kFirstBBlock + 2 * 768 * 4 + 64 * 4 + 16 * 8 + 16;

__global__ void __launch_bounds__(kWsThreads, 3) conv_first_ws_kernel(const FirstParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sA = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // [2][3 x 4 KB]
    uint8_t* sB = sA + 2 * 3 * kFirstABlock;                                        // [3 x 2 KB]
    float* s_in = reinterpret_cast<float*>(sB + 3 * kFirstBBlock);                  // [2][768]
    float* s_bias = s_in + 2 * 768;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 64);                      // a_full[2] a_empty[2] t_full[2] t_empty[2] pbar
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, W = p.W, cin = p.cin, cout = p.cout, tiles_x = p.tiles_x, tiles_y = p.tiles_y;
    int* const err = p.err;
    const int tmem_cols = cout <= 16 ? 32 : (cout <= 32 ? 64 : 128);                // two accumulators

    for (int i = tid; i < 3 * kFirstBBlock / 16; i += kWsThreads) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * kFirstABlock / 16; i += kWsThreads) {                 // K block 2 of both A buffers: taps 9..11 stay zero
        const int b = i / (kFirstABlock / 16), r = i - b * (kFirstABlock / 16);
        reinterpret_cast<uint4*>(sA + b * 3 * kFirstABlock + 2 * kFirstABlock)[r] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (int i = tid; i < cout * 36; i += kWsThreads) {
        const int nrow = i / 36, k = i - nrow * 36, c = k & 3, tap = k >> 2;
        if (c < cin) {
            const int kb = k >> 4, kk = k & 15;
            *reinterpret_cast<__nv_bfloat16*>(sB + kb * kFirstBBlock + swz32((uint32_t)nrow, (uint32_t)(kk >> 3)) + (kk & 7) * 2) =
                __float2bfloat16_rn(p.w[(nrow * cin + c) * 9 + tap]);
        }
    }
    for (int i = tid; i < 64; i += kWsThreads) s_bias[i] = (i < cout && p.bias) ? p.bias[i] : 0.f;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars[0 + b]), 128); mbar_init(smem_u32(&bars[2 + b]), 1);
            mbar_init(smem_u32(&bars[4 + b]), 1);   mbar_init(smem_u32(&bars[6 + b]), 4);
        }
        mbar_init(smem_u32(&bars[8]), 128);
        fence_mbarrier_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(s_tmem), (uint32_t)tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[2]), t_full = smem_u32(&bars[4]), t_empty = smem_u32(&bars[6]),
                   pbar = smem_u32(&bars[8]);
    const int total = p.n * tiles_y * tiles_x;
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_begin = min(total, (int)blockIdx.x * per), t_end = min(total, t_begin + per);
    int t_img = t_begin / (tiles_y * tiles_x), t_ty, t_tx;
    { const int r = t_begin - t_img * (tiles_y * tiles_x); t_ty = r / tiles_x; t_tx = r - t_ty * tiles_x; }

    if (warp < 4) {
        // ============================== producers ==============================
        const float* const in = p.in;
        const size_t plane = (size_t)H * W;
        int e_off[6], e_row[6], e_col[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int idx = lane + 32 * j;
            e_row[j] = idx / kFirstHaloW;
            e_col[j] = idx - e_row[j] * kFirstHaloW;
            e_off[j] = e_row[j] * W + e_col[j];
        }
        const bool ch_on = warp < cin, last_on = lane + 160 < kFirstPlane;
        float pre[6];
        auto prefetch = [&]() {
            const int y0 = t_ty * kFirstTileH - 1, x0 = t_tx * kFirstTileW - 1;
            const float* base = in + ((size_t)t_img * cin + warp) * plane + (long long)y0 * W + x0;
            if (!ch_on) {
#pragma unroll
                for (int j = 0; j < 6; ++j) pre[j] = 0.f;
            } else if (y0 >= 0 && x0 >= 0 && y0 + kFirstHaloH <= H && x0 + kFirstHaloW <= W) {
#pragma unroll
                for (int j = 0; j < 5; ++j) pre[j] = __ldg(base + e_off[j]);
                pre[5] = last_on ? __ldg(base + e_off[5]) : 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int gy = y0 + e_row[j], gx = x0 + e_col[j];
                    const bool ok = (j < 5 || last_on) && gy >= 0 && gy < H && gx >= 0 && gx < W;
                    pre[j] = ok ? __ldg(base + e_off[j]) : 0.f;
                }
            }
            if (++t_tx == tiles_x) { t_tx = 0; if (++t_ty == tiles_y) { t_ty = 0; ++t_img; } }
        };
        if (t_begin < t_end) prefetch();
        const int my = tid >> 4, mx = tid & 15;
        for (int t = t_begin, k = 0; t < t_end; ++t, ++k) {
            const int b = k & 1;
            // staging tile: bf16, channel-interleaved ([pixel][4 channels] = 8 bytes per pixel), converted by the loading thread —
            // the im2col below then reads a tap's four channels with ONE 8-byte load that already is the packed pair of words of
            // the A row (36 scalar loads + 18 conversions per pixel before: the kernel is bound by load / store-unit work)
            uint16_t* const s_mine = reinterpret_cast<uint16_t*>(s_in + b * 768) + lane * 4 + warp;
#pragma unroll
            for (int j = 0; j < 5; ++j) s_mine[128 * j] = bf1(pre[j]);
            if (last_on) s_mine[640] = bf1(pre[5]);
            mbar_arrive(pbar);                                               // producer-group barrier: the fp32 tile is complete
            if (t + 1 < t_end) prefetch();                                   // next tile's loads fly during the im2col and the waits
            mbar_wait(pbar, (uint32_t)(k & 1), err, 311);
            mbar_wait(a_empty + 8 * b, (uint32_t)(((k >> 1) & 1) ^ 1), err, 312);      // the MMAs that read A[b] two tiles ago are done
            const uint2* const q0 = reinterpret_cast<const uint2*>(s_in + b * 768) + my * kFirstHaloW + mx;
            uint8_t* const a = sA + b * 3 * kFirstABlock;
            uint32_t pk[9][2];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const uint2 q = q0[ky * kFirstHaloW + kx];               // channels (0, 1) | (2, 3) of the tap's pixel
                    pk[ky * 3 + kx][0] = q.x;
                    pk[ky * 3 + kx][1] = q.y;
                }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
                *reinterpret_cast<uint4*>(a + (ch >> 1) * kFirstABlock + swz32((uint32_t)tid, (uint32_t)(ch & 1))) =
                    make_uint4(pk[2 * ch][0], pk[2 * ch][1], pk[2 * ch + 1][0], pk[2 * ch + 1][1]);
            *reinterpret_cast<uint4*>(a + 2 * kFirstABlock + swz32((uint32_t)tid, 0u)) = make_uint4(pk[8][0], pk[8][1], 0u, 0u);
            fence_proxy_async();
            mbar_arrive(a_full + 8 * b);
        }
    } else if (warp == 8) {
        // ============================== MMA issuer ==============================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dhi = umma_desc_hi(32);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int t = t_begin, k = 0; t < t_end; ++t, ++k) {
            const int b = k & 1;
            const uint32_t par = (uint32_t)((k >> 1) & 1);
            mbar_wait(t_empty + 8 * b, par ^ 1, err, 313);                   // the epilogue has drained accumulator b
            mbar_wait(a_full + 8 * b, par, err, 314);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 3; ++kb)
                    tc_mma_bf16(tmem_base + (uint32_t)(b * cout), umma_desc(dhi, a0 + (b * 3 + kb) * kFirstABlock),
                                umma_desc(dhi, b0 + kb * kFirstBBlock), idesc, kb ? 1u : 0u);
                tc_commit(a_empty + 8 * b);
                tc_commit(t_full + 8 * b);
            }
            __syncwarp();
        }
    } else {
        // ============================== epilogue (warps 4..7) ==============================
        const float slope = p.slope;
        __nv_bfloat16* const out = p.out;
        const int e = tid - 128, my = e >> 4, mx = e & 15;                   // accumulator row == pixel inside the tile
        const uint32_t taddr0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint64_t slope2 = f2_pack(slope, slope), zero2 = f2_pack(0.0f, 0.0f);
        for (int t = t_begin, k = 0; t < t_end; ++t, ++k) {
            const int b = k & 1;
            const int y = t_ty * kFirstTileH + my, x = t_tx * kFirstTileW + mx;
            const bool valid = y < H && x < W;
            __nv_bfloat16* op = out + (((size_t)t_img * H + y) * (size_t)W + x) * cout;
            if (++t_tx == tiles_x) { t_tx = 0; if (++t_ty == tiles_y) { t_ty = 0; ++t_img; } }
            mbar_wait(t_full + 8 * b, (uint32_t)((k >> 1) & 1), err, 315);
            tc_fence_after();
            const uint32_t taddr = taddr0 + (uint32_t)(b * cout);
            for (int j = 0; j < cout / 16; j += 2) {                         // two 16-column chunks per TMEM wait
                uint32_t v[2][16];
                const int nj = min(2, cout / 16 - j);
                tc_ld16(taddr + j * 16, v[0]);
                if (nj > 1) tc_ld16(taddr + (j + 1) * 16, v[1]);
                tc_ld_wait();
                if (j + 2 >= cout / 16) {                                    // all columns read: hand the accumulator back before the math
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty + 8 * b);
                }
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    if (h2 < nj) {
                        uint32_t o[8];
                        const float4* b4 = reinterpret_cast<const float4*>(s_bias + (j + h2) * 16);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 bb = b4[i];
                            const uint64_t a01 = f2_add(f2_pack(__uint_as_float(v[h2][4 * i]), __uint_as_float(v[h2][4 * i + 1])), f2_pack(bb.x, bb.y));
                            const uint64_t a23 = f2_add(f2_pack(__uint_as_float(v[h2][4 * i + 2]), __uint_as_float(v[h2][4 * i + 3])), f2_pack(bb.z, bb.w));
                            float a0_, a1_, a2_, a3_, m0, m1, m2, m3;
                            f2_unpack(a01, a0_, a1_); f2_unpack(a23, a2_, a3_);
                            f2_unpack(f2_fma(a01, slope2, zero2), m0, m1); f2_unpack(f2_fma(a23, slope2, zero2), m2, m3);
                            o[2 * i] = bf2(fmaxf(a0_, m0), fmaxf(a1_, m1));
                            o[2 * i + 1] = bf2(fmaxf(a2_, m2), fmaxf(a3_, m3));
                        }
                        if (valid) st_global_256(op + (j + h2) * 16, o);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)tmem_cols); }
}

static int* g_first_err = nullptr;

}  // namespace pnnp

using namespace pnnp;

#ifdef PNNP_HOST_EMUL
#define PNNP_FIRST_KLAUNCH(grid) emul_launch_1d(grid, kFirstThreads, [&]() { conv_first_kernel(p); })
#define PNNP_FIRST_WS_KLAUNCH(grid) emul_launch_1d(grid, kWsThreads, [&]() { conv_first_ws_kernel(p); })
#else
#define PNNP_FIRST_KLAUNCH(grid) conv_first_kernel<<<grid, kFirstThreads, kFirstSmem, (cudaStream_t)stream>>>(p)
#define PNNP_FIRST_WS_KLAUNCH(grid) conv_first_ws_kernel<<<grid, kWsThreads, kWsSmem, (cudaStream_t)stream>>>(p)
#endif

extern "C" int pnnp_conv_first_nchw(const float* in, const float* weight, const float* bias, void* out, int n, int cin, int h, int w,
                                    int cout, int act, void* stream) {
    if (!in || !weight || !out) return fail("conv_first: null pointer");
    if (cin < 1 || cin > 4) return fail("conv_first: 1..4 input channels (a packed Bayer frame)");
    if (cout % 16 || cout < 16 || cout > 64) return fail("conv_first: cout must be 16, 32, 48 or 64");
    if (n < 1 || h < 1 || w < 1) return 0;
    FirstParams p{};
    p.in = in; p.w = weight; p.bias = bias; p.out = static_cast<__nv_bfloat16*>(out);
    p.n = n; p.cin = cin; p.H = h; p.W = w; p.cout = cout;
    p.tiles_x = (w + kFirstTileW - 1) / kFirstTileW; p.tiles_y = (h + kFirstTileH - 1) / kFirstTileH;
    p.slope = act == 1 ? 0.2f : (act == 2 ? 0.f : 1.f);                      // ACT_LEAKY / ACT_RELU / ACT_NONE of the conv kernels
    if (!g_first_err) { PNNP_CUDA(cudaMalloc(&g_first_err, sizeof(int))); PNNP_CUDA(cudaMemset(g_first_err, 0, sizeof(int))); }
    p.err = g_first_err;
    p.dbg = getenv("PNNP_FIRST_DBG") ? atoi(getenv("PNNP_FIRST_DBG")) : 0;
    int dev = 0, sms = 0;
    PNNP_CUDA(cudaGetDevice(&dev));
    PNNP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long total = (long long)n * p.tiles_y * p.tiles_x;
    if (total > 0x7FFFFFFFll) return fail("conv_first: too many tiles");
    if (variant_on("PNNP_FIRST_WS") && !p.dbg) {                           // warp-specialised form (default); =0: the single-role kernel
#ifndef PNNP_HOST_EMUL
        static bool attr_done = false;
        if (!attr_done) { PNNP_CUDA(cudaFuncSetAttribute(conv_first_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmem)); attr_done = true; }
#endif
        const int grid_ws = (int)std::min<long long>(total, (long long)sms * 3);
        PNNP_FIRST_WS_KLAUNCH(grid_ws);
    } else {
        const int grid = (int)std::min<long long>(total, (long long)sms * 8);
        PNNP_FIRST_KLAUNCH(grid);
    }
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_conv_first_pipeline_error(void) {
    int v = 0;
    if (g_first_err) { cudaMemcpy(&v, g_first_err, sizeof(int), cudaMemcpyDeviceToHost); if (v) cudaMemset(g_first_err, 0, sizeof(int)); }
    return v;
}
