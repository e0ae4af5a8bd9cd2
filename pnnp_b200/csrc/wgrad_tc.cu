// Weight-gradient GEMM on tcgen05/TMEM for the training step (T1).
//
//   dW[tap][co][ci] = sum over pixels q of  G^T[co][q] * X^T[ci][q + off(tap)]
//
// G^T and X^T are the pre-activation gradient and the layer input stored channel-major with a zero
// ring around every image (transpose_pad_kernel), so a filter tap is a plain offset off(tap) =
// (dy-1)*wp + (dx-1) in the flattened padded pixel index.  A TMA box must start at a 16-byte aligned
// innermost coordinate (an odd pixel offset raises an illegal-instruction fault), so the row pitch wp is
// a multiple of 8 pixels and X^T is stored in three x-shifted planes (dx = 0,1,2): tap (dy,dx) reads plane
// dx at offset (dy-1)*wp.  Both operands are then ordinary K-major
// GEMM tiles (K = pixels contiguous, 128-byte swizzle): A = 128 rows of G^T x 64 pixels, B = the rows of
// X^T for up to `taps` taps stacked in shared memory (tap t at row block t*ci_tile), so one MMA covers
// N = up to 256 columns = several taps.  Accumulators for all taps of the pass live in TMEM
// (taps * ci_tile <= 512 columns).  Split-K: each CTA owns a range of 64-pixel steps and adds its partial
// tile to the fp32 gradient with red.global.add (order-dependent fp32 summation, as in any split-K wgrad).
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (run once, after the K loop).
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "abi_common.h"
#include "tc_common.cuh"
#include "../../include/pnnp_b200.h"

namespace pnnp {

constexpr int kWgBM = 128, kWgBK = 64, kWgMaxTaps = 9, kWgStagesMax = 6;

struct WgradParams {
    int taps, ci_tile, ncols;       // accumulator columns = taps * ci_tile
    int m_tiles, n_tiles, splits, ksteps_total;
    int stages, stage_bytes, b_bytes_per_tap;
    int tap_off[kWgMaxTaps];        // pixel offset of each tap of this pass (multiple of 8: TMA start alignment)
    int tap_plane[kWgMaxTaps];      // which x-shifted copy of X^T the tap reads
    int tap_id[kWgMaxTaps];         // index of each tap in the output layout
    int co, ci, ci_off, ci_total;   // rows valid in A, columns valid per tap, column offset / stride of dW
    float* dw;                      // [tap][co][ci_total] fp32
    int tmem_cols;
    int dbg;                        // PNNP_WG_DBG bits: 1 skip TMA, 2 skip MMA, 4 skip TMEM loads (pipeline debugging)
    int* err;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWgStagesMax;
    uint64_t* done_bar = bars + 2 * kWgStagesMax;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStagesMax + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item: (split, m_tile, n_tile)
    int r = blockIdx.x;
    const int split = r % p.splits; r /= p.splits;
    const int n_tile = r % p.n_tiles; r /= p.n_tiles;
    const int m_tile = r;
    const int per = (p.ksteps_total + p.splits - 1) / p.splits;
    const int k_begin = split * per, k_end = min(p.ksteps_total, k_begin + per);
    const int nk = max(0, k_end - k_begin);
    const uint32_t a_bytes = kWgBM * kWgBK * 2;
    const uint32_t stage_tx = a_bytes + (uint32_t)(p.taps * p.ci_tile * kWgBK * 2);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        uint32_t stage = 0, phase = 0;
        for (int ks = 0; ks < nk; ++ks) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.err, 201);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
            const int q0 = (k_begin + ks) * kWgBK;
            if (p.dbg & 1) { if (elect_one()) mbar_arrive(fb); }
            else if (elect_one()) {
                mbar_expect_tx(fb, stage_tx);
                tma_load_2d(sa, &tmA, fb, q0, m_tile * kWgBM);
                for (int t = 0; t < p.taps; ++t)
                    tma_load_3d(sa + a_bytes + t * p.b_bytes_per_tap, &tmB, fb, q0 + p.tap_off[t], n_tile * p.ci_tile, p.tap_plane[t]);
            }
            __syncwarp();
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        const uint64_t dhi = umma_desc_hi(128);
        uint32_t stage = 0, phase = 0;
        for (int ks = 0; ks < nk; ++ks) {
            mbar_wait(smem_u32(&full_bar[stage]), phase, p.err, 203);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
            const uint64_t adesc0 = dhi | (uint64_t)((sa >> 4) & 0x3FFFu);
            if (elect_one()) {
                // stacked taps: columns [c0, c0 + n) of the accumulator <- B rows [c0, c0 + n)
                for (int c0 = 0; c0 < p.ncols; c0 += 256) {
                    const int n = min(256, p.ncols - c0);
                    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kWgBM >> 4) << 24);
                    const uint32_t sb = sa + a_bytes + (uint32_t)c0 * (kWgBK * 2);
                    const uint64_t bdesc0 = dhi | (uint64_t)((sb >> 4) & 0x3FFFu);
#pragma unroll
                    for (int k = 0; k < kWgBK / 16; ++k)
                        if (!(p.dbg & 2)) tc_mma_bf16(tmem_base + (uint32_t)c0, adesc0 + (uint64_t)(k * 2), bdesc0 + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
                }
                tc_commit(smem_u32(&empty_bar[stage]));
            }
            __syncwarp();
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit(smem_u32(done_bar));
        __syncwarp();
    } else if (nk > 0) {
        // epilogue: thread = accumulator row = output channel co; add the partial tile to dW
        const int quad = warp & 3;
        const int row = m_tile * kWgBM + quad * 32 + lane;
        mbar_wait(smem_u32(done_bar), 0, p.err, 204);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int t = 0; t < p.taps; ++t) {
            float* dst_row = p.dw + ((size_t)p.tap_id[t] * p.co + row) * p.ci_total + p.ci_off + n_tile * p.ci_tile;
            for (int j = 0; j < p.ci_tile; j += 16) {
                uint32_t v[16];
                if (p.dbg & 4) continue;
                tc_ld16(taddr + t * p.ci_tile + j, v);
                tc_ld_wait();
                if (row < p.co && !(p.dbg & 8)) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n_tile * p.ci_tile + j + i < p.ci) atomicAdd(dst_row + j + i, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
    }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 get_encode2() {
    static EncodeTiledFn2 fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn2>(q);
    }
    return fn;
}
// channel-major operand [planes][rows][row_elems] bf16 viewed as (pixels, rows[, planes]); box (64 pixels, box_rows[, 1])
static int make_cm_map(CUtensorMap* tm, const void* ptr, size_t row_elems, int rows, int box_rows, int planes) {
    EncodeTiledFn2 enc = get_encode2();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)row_elems, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)row_elems * 2, (cuuint64_t)row_elems * 2 * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)kWgBK, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, planes > 0 ? 3 : 2, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { char b[128]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled(wgrad operand) failed: %d", (int)r); return fail(b); }
    return 0;
}

static int* g_wg_err = nullptr;

}  // namespace pnnp

using namespace pnnp;

// gT: [co][row_elems], xT: [planes][ci][row_elems] (bf16, channel-major, zero ring); tap t reads plane tap_plane_host[t]
// at pixel offset tap_off_host[t]; dw: fp32 [taps_total][co][ci_total], columns [ci_off, ci_off + ci) are accumulated.
extern "C" int pnnp_wgrad_tc(const void* gT, const void* xT, size_t row_elems, size_t valid_elems, int co, int ci, int taps_total,
                             const int* tap_off_host, const int* tap_plane_host, int planes, float* dw, int ci_off, int ci_total,
                             void* stream) {
    if (!gT || !xT || !dw || !tap_off_host) return fail("wgrad: null pointer");
    if (planes < 1) return fail("wgrad: planes must be >= 1");
    for (int t = 0; t < taps_total && t < kWgMaxTaps; ++t) {
        if (tap_off_host[t] % 8) return fail("wgrad: tap offsets must be multiples of 8 pixels (16-byte aligned TMA start)");
        if (tap_plane_host && (tap_plane_host[t] < 0 || tap_plane_host[t] >= planes)) return fail("wgrad: tap plane out of range");
    }
    if ((row_elems % 8) || taps_total < 1 || taps_total > kWgMaxTaps || (ci % 8) || co < 1) return fail("wgrad: bad shape (row pitch and ci must be multiples of 8)");
    cudaStream_t st = (cudaStream_t)stream;
    const int ci_tile = std::min(ci, 256);
    if (ci % ci_tile) return fail("wgrad: ci must be <= 256 or a multiple of 256");
    const int n_tiles = ci / ci_tile, m_tiles = (co + kWgBM - 1) / kWgBM;
    // taps per pass: accumulator columns <= 512 and the stacked B tile <= 48 KB per stage
    int tpp = std::min(taps_total, std::min(512 / ci_tile, 384 / ci_tile));
    if (tpp < 1) tpp = 1;
    const int ksteps_total = (int)((valid_elems + kWgBK - 1) / kWgBK);
    int dev = 0, sms = 0;
    PNNP_CUDA(cudaGetDevice(&dev));
    PNNP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!g_wg_err) { PNNP_CUDA(cudaMalloc(&g_wg_err, sizeof(int))); PNNP_CUDA(cudaMemset(g_wg_err, 0, sizeof(int))); }
    CUtensorMap tmA, tmB;
    if (int e = make_cm_map(&tmA, gT, row_elems, co, kWgBM, 0)) return e;
    if (int e = make_cm_map(&tmB, xT, row_elems, ci, ci_tile, planes)) return e;
    static bool attr_done = false;
    if (!attr_done) { PNNP_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_done = true; }
    for (int t0 = 0; t0 < taps_total; t0 += tpp) {
        WgradParams p{};
        p.taps = std::min(tpp, taps_total - t0); p.ci_tile = ci_tile; p.ncols = p.taps * ci_tile;
        p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.ksteps_total = ksteps_total;
        const int combos = m_tiles * n_tiles;
        p.splits = std::max(1, std::min(ksteps_total, (2 * sms + combos - 1) / combos));
        p.b_bytes_per_tap = ci_tile * kWgBK * 2;
        p.stage_bytes = (kWgBM * kWgBK * 2 + p.taps * p.b_bytes_per_tap + 1023) / 1024 * 1024;
        p.stages = std::max(2, std::min(kWgStagesMax, (227 * 1024 - 2048) / p.stage_bytes));
        for (int t = 0; t < p.taps; ++t) { p.tap_off[t] = tap_off_host[t0 + t]; p.tap_plane[t] = tap_plane_host ? tap_plane_host[t0 + t] : 0; p.tap_id[t] = t0 + t; }
        p.co = co; p.ci = ci; p.ci_off = ci_off; p.ci_total = ci_total; p.dw = dw;
        int tc = 32; while (tc < p.ncols) tc <<= 1;
        p.tmem_cols = tc; p.err = g_wg_err;
        { const char* e = getenv("PNNP_WG_DBG"); p.dbg = e ? atoi(e) : 0; }
        const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + (2 * kWgStagesMax + 1) * 8 + 64;
        if (smem > 227 * 1024) return fail("wgrad: shared memory budget exceeded");
        wgrad_tc_kernel<<<combos * p.splits, 192, smem, st>>>(tmA, tmB, p);
        count_launch();
        PNNP_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int pnnp_wgrad_pipeline_error(void) {
    int v = 0;
    if (g_wg_err) { cudaMemcpy(&v, g_wg_err, sizeof(int), cudaMemcpyDeviceToHost); if (v) cudaMemset(g_wg_err, 0, sizeof(int)); }
    return v;
}
