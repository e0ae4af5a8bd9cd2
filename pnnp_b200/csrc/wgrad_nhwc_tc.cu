// Weight gradients of the training step (T1) as an implicit GEMM straight from the NHWC bf16 activations — no
// transposed copies, no padding pass:
//
//   3x3 conv      dW[ky*3+kx][ci][co] = sum over pixels p of  x[p + (ky-1, kx-1)][ci] * g[p][co]        (zero outside the image)
//   convT 2x2 s2  dW[a*2+b][ci][co]   = sum over input pixels p of  x[p][ci] * g[2p + (a, b)][co]
//
// GEMM view: K = pixels, M = output channels of the layer (rows of g), N = (filter tap, input channel) (rows of x).  A TMA box
// of an NHWC tensor — `cw` channels x 16 x 8 pixels (10 rows for the 3x3 halo) — lands in shared memory as [pixel][cw channels],
// i.e. K rows of MN-contiguous elements: the canonical **MN-major** operand layout of tcgen05.mma (instruction-descriptor bits
// 15/16 set, 32/64/128-byte swizzle = the row pitch, SBO = 8 rows, LBO = distance between channel blocks).  The conv's zero
// padding is the TMA unit's out-of-bounds fill; a shift by one filter row is a 16-row offset of the operand start inside the
// haloed box; a shift by one filter column is a separate box; the stride-2 gather of the transposed conv is the tensor map's
// element stride.  One MMA covers 128 output channels x up to 192 (tap, channel) columns x 16 pixels; accumulators stay in
// TMEM (<= 512 columns) over the CTA's whole pixel range (split-K across CTAs) and are added to the fp32 gradient with
// coalesced red.global.add (the scratch layout is [tap][ci][co], output channel = TMEM lane = fastest index).
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (once, after the K loop).
#ifndef PNNP_HOST_EMUL
#include <cuda.h>
#include <cuda_bf16.h>
#endif
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "abi_common.h"
#include "tc_common.cuh"
#include "../../include/pnnp_b200.h"

// kernel launch; tests/emul/ compiles this file for the host and runs the launch on its SIMT emulator + tensor-core model instead
#ifdef PNNP_HOST_EMUL
#define PNNP_WGRAD_KLAUNCH(V) emul_launch_1d(combos * splits, 192, [&]() { wgrad_nhwc_kernel<V>(tmG, tmX, p); })
#else
#define PNNP_WGRAD_KLAUNCH(V) wgrad_nhwc_kernel<V><<<combos * splits, 192, smem, st>>>(tmG, tmX, p)
#endif

namespace pnnp {

constexpr int kWnTileW = 16, kWnTileH = 8, kWnPix = 128;      // pixels per K stage
constexpr int kWnMaxBoxes = 6, kWnMaxMmas = 3, kWnMaxAtoms = 9, kWnStagesMax = 4;

// coords (c0, X0*mul+dx, Y0*mul+dy, img); by_ts: 1 = filter column from the CTA's tap set (dx = ts - 1), 2 = transposed-conv tap
// (dx, dy) = (ts & 1, ts >> 1); c0 advances with the CTA's M tile (g boxes, 128 channels) or N tile (x boxes, ci_tile channels)
struct WnBox { int is_x; int c0; int mul; int dx, dy; int by_ts; int smem_off; int bytes; };
struct WnMma { int a_off, b_off; int n; int lbo_b; int tmem_col; };                  // one MMA per 16-pixel K step
struct WnAtom { int col, width, tap, ci0, ky_base; };   // rows-by-filter-row mode: row block rb holds ky = ky_base - rb (valid 0..2), tap = kx                                        // accumulator columns -> (tap [+ ts], input channels [+ N tile])

struct WnParams {
    int n_boxes, n_mmas, n_atoms;
    WnBox box[kWnMaxBoxes];
    WnMma mma[kWnMaxMmas];
    WnAtom atom[kWnMaxAtoms];
    int pitch_a, pitch_b, lbo_a;       // bytes per pixel row of the g / x tiles (= swizzle span); distance between g channel blocks
    int swz_a, swz_b;
    int stage_bytes, stage_tx, stages;
    int tiles_x, tiles_y, tiles_total; // pixel tiles of the K range (per image tiles_x * tiles_y)
    int splits, tapsets, n_tiles, ci_tile, tap_by_ts;   // grid = m_tiles * n_tiles * tapsets * splits
    int co;                            // valid accumulator rows of M tile mt: co - 128 * mt
    int row_block;                     // != 0: accumulator rows are (filter row, co) in blocks of row_block rows, see WnAtom::ky_base
                                       // (3x3 layers with 32 / 64 output channels)
    int ci_total, co_pad;              // dw scratch geometry [tap][ci_total][co_pad]
    float* dw;
    int tmem_cols;
    int dbg;
    int* err;
};

// MN-major shared-memory matrix descriptor: [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) swizzle
__device__ __forceinline__ uint64_t umma_desc_mn_hi(int swz, int lbo_bytes) {
    const uint64_t layout = swz == 128 ? 2ull : (swz == 64 ? 4ull : 6ull);
    const uint64_t sbo = (uint64_t)((8 * swz) >> 4);                 // 8 pixel rows of `swz` bytes
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// VAR 1 (the default since r02: 4.88 -> 4.82 ms per step; PNNP_WGRAD_V2=0 for the first form): the producer and MMA warps — one warp each, so the instruction count of their
// per-stage loops is the stage rate (352 and 139 SASS instructions in VAR 0: two integer divisions and a walk over the box / MMA
// tables in the kernel-parameter bank per stage) — keep everything that does not change from stage to stage in registers: the
// boxes' coordinates offsets, shared-memory offsets and tensor maps, the MMAs' descriptors, the (image, tile row, tile column) of the
// stage advanced by increment-and-carry, the stage's barrier / buffer addresses advanced by a constant.  Same loads, same MMAs.
template <int VAR>
__global__ void __launch_bounds__(192, 1)
wgrad_nhwc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by POINTER arithmetic on the __shared__ array (not through an integer cast): the compiler keeps the
    // shared address space of everything carved from it, so plain loads / stores of these pointers are LDS / STS instead of
    // generic LD / ST (which go through the global-memory instruction queue: stall reason lg_throttle in the r02 capture)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWnStagesMax;
    uint64_t* done_bar = bars + 2 * kWnStagesMax;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWnStagesMax + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bid = blockIdx.x;
    const int split = bid % p.splits; bid /= p.splits;
    const int ts = bid % p.tapsets; bid /= p.tapsets;
    const int nt = bid % p.n_tiles;
    const int mt = bid / p.n_tiles;
    const int per = (p.tiles_total + p.splits - 1) / p.splits;
    const int t_begin = split * per, t_end = min(p.tiles_total, t_begin + per);
    const int nk = max(0, t_end - t_begin);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmG);
        prefetch_tensormap(&tmX);
        for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(done_bar), 1);
        fence_mbarrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (VAR == 1 && warp == 0) {
        // ============================== TMA producer, loop-invariant state in registers ==============================
        int bc0[kWnMaxBoxes], bdx[kWnMaxBoxes], bdy[kWnMaxBoxes], bmul[kWnMaxBoxes], boff[kWnMaxBoxes];
        bool bisx[kWnMaxBoxes];
#pragma unroll
        for (int b = 0; b < kWnMaxBoxes; ++b) {
            const WnBox& bx = p.box[b < p.n_boxes ? b : 0];
            bisx[b] = bx.is_x != 0;
            bc0[b] = bx.c0 + (bx.is_x ? nt * p.ci_tile : mt * 128);
            bdx[b] = bx.by_ts == 1 ? ts - 1 : (bx.by_ts == 2 ? (ts & 1) : bx.dx);
            bdy[b] = bx.by_ts == 2 ? (ts >> 1) : bx.dy;
            bmul[b] = bx.mul; boff[b] = bx.smem_off;
        }
        const int n_boxes = p.n_boxes, tiles_x = p.tiles_x, tiles_y = p.tiles_y, stages = p.stages, stage_bytes = p.stage_bytes;
        const uint32_t stage_tx = (uint32_t)p.stage_tx, smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        int* const err = p.err;
        const int tiles_img = tiles_x * tiles_y;
        int img = t_begin / tiles_img, ty = (t_begin - img * tiles_img) / tiles_x, tx = t_begin - img * tiles_img - ty * tiles_x;
        uint32_t stage = 0, phase = 0, sa = smem0, fb = full0, eb = empty0;
        for (int t = t_begin; t < t_end; ++t) {
            const int X0 = tx * kWnTileW, Y0 = ty * kWnTileH;
            mbar_wait(eb, phase ^ 1, err, 301);
            if (elect_one()) {
                mbar_expect_tx(fb, stage_tx);
#pragma unroll
                for (int b = 0; b < kWnMaxBoxes; ++b)
                    if (b < n_boxes)
                        tma_load_4d(sa + (uint32_t)boff[b], bisx[b] ? &tmX : &tmG, fb, bc0[b], X0 * bmul[b] + bdx[b], Y0 * bmul[b] + bdy[b], img);
            }
            __syncwarp();
            if (++tx == tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++img; } }
            sa += (uint32_t)stage_bytes; fb += 8; eb += 8;
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; sa = smem0; fb = full0; eb = empty0; }
        }
    } else if (VAR == 1 && warp == 1) {
        // ============================== MMA issuer, descriptors in registers ==============================
        const uint64_t ahi = umma_desc_mn_hi(p.swz_a, p.lbo_a);
        const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t kstep_a = (uint32_t)(16 * p.pitch_a) >> 4, kstep_b = (uint32_t)(16 * p.pitch_b) >> 4;
        uint32_t m_idesc[kWnMaxMmas], m_aoff[kWnMaxMmas], m_boff[kWnMaxMmas], m_col[kWnMaxMmas];
        uint64_t m_bhi[kWnMaxMmas];
#pragma unroll
        for (int m = 0; m < kWnMaxMmas; ++m) {
            const WnMma& mm = p.mma[m < p.n_mmas ? m : 0];
            m_idesc[m] = idesc_base | ((uint32_t)(mm.n >> 3) << 17);
            m_aoff[m] = (uint32_t)mm.a_off; m_boff[m] = (uint32_t)mm.b_off; m_col[m] = tmem_base + (uint32_t)mm.tmem_col;
            m_bhi[m] = umma_desc_mn_hi(p.swz_b, mm.lbo_b);
        }
        const int n_mmas = p.n_mmas, stages = p.stages, stage_bytes = p.stage_bytes;
        const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        int* const err = p.err;
        uint32_t stage = 0, phase = 0, sa = smem0, fb = full0, eb = empty0;
        for (int ks = 0; ks < nk; ++ks) {
            mbar_wait(fb, phase, err, 303);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int m = 0; m < kWnMaxMmas; ++m) {
                    if (m < n_mmas) {
                        const uint64_t adesc0 = ahi | (uint64_t)(((sa + m_aoff[m]) >> 4) & 0x3FFFu);
                        const uint64_t bdesc0 = m_bhi[m] | (uint64_t)(((sa + m_boff[m]) >> 4) & 0x3FFFu);
#pragma unroll
                        for (int k = 0; k < kWnPix / 16; ++k)
                            tc_mma_bf16(m_col[m], adesc0 + (uint64_t)(k * kstep_a), bdesc0 + (uint64_t)(k * kstep_b), m_idesc[m], (ks | k) != 0 ? 1u : 0u);
                    }
                }
                tc_commit(eb);
            }
            __syncwarp();
            sa += (uint32_t)stage_bytes; fb += 8; eb += 8;
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; sa = smem0; fb = full0; eb = empty0; }
        }
        if (nk > 0 && elect_one()) tc_commit(smem_u32(done_bar));      // an empty K range has no epilogue: nobody would wait for this commit
        __syncwarp();
    } else if (warp == 0) {
        // ============================== TMA producer ==============================
        uint32_t stage = 0, phase = 0;
        const int tiles_img = p.tiles_x * p.tiles_y;
        for (int t = t_begin; t < t_end; ++t) {
            const int img = t / tiles_img, r = t - img * tiles_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            const int X0 = tx * kWnTileW, Y0 = ty * kWnTileH;
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.err, 301);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
            if (p.dbg & 1) { if (elect_one()) mbar_arrive(fb); }
            else if (elect_one()) {
                mbar_expect_tx(fb, (uint32_t)p.stage_tx);
                for (int b = 0; b < p.n_boxes; ++b) {
                    const WnBox& bx = p.box[b];
                    const int c0 = bx.c0 + (bx.is_x ? nt * p.ci_tile : mt * 128);
                    const int dx = bx.by_ts == 1 ? ts - 1 : (bx.by_ts == 2 ? (ts & 1) : bx.dx);
                    const int dy = bx.by_ts == 2 ? (ts >> 1) : bx.dy;
                    tma_load_4d(sa + bx.smem_off, bx.is_x ? &tmX : &tmG, fb, c0, X0 * bx.mul + dx, Y0 * bx.mul + dy, img);
                }
            }
            __syncwarp();
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        const uint64_t ahi = umma_desc_mn_hi(p.swz_a, p.lbo_a);
        // instruction descriptor: D fp32, A/B bf16, both operands MN-major (bits 15, 16), M = 128
        const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t kstep_a = (uint32_t)(16 * p.pitch_a) >> 4, kstep_b = (uint32_t)(16 * p.pitch_b) >> 4;
        uint32_t stage = 0, phase = 0;
        for (int ks = 0; ks < nk; ++ks) {
            mbar_wait(smem_u32(&full_bar[stage]), phase, p.err, 303);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
            if (elect_one()) {
                for (int m = 0; m < p.n_mmas; ++m) {
                    const WnMma& mm = p.mma[m];
                    const uint32_t idesc = idesc_base | ((uint32_t)(mm.n >> 3) << 17);
                    const uint64_t adesc0 = ahi | (uint64_t)(((sa + (uint32_t)mm.a_off) >> 4) & 0x3FFFu);
                    const uint64_t bdesc0 = umma_desc_mn_hi(p.swz_b, mm.lbo_b) | (uint64_t)(((sa + (uint32_t)mm.b_off) >> 4) & 0x3FFFu);
#pragma unroll
                    for (int k = 0; k < kWnPix / 16; ++k)
                        if (!(p.dbg & 2))
                            tc_mma_bf16(tmem_base + (uint32_t)mm.tmem_col, adesc0 + (uint64_t)(k * kstep_a), bdesc0 + (uint64_t)(k * kstep_b),
                                        idesc, (ks | k) != 0 ? 1u : 0u);
                }
                tc_commit(smem_u32(&empty_bar[stage]));
            }
            __syncwarp();
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit(smem_u32(done_bar));
        __syncwarp();
    } else if (nk > 0) {
        // ============================== epilogue ==============================
        // thread = accumulator row = output channel; add the partial tile to dw[tap][ci][co] (co fastest: coalesced reds)
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int m0 = mt * 128;
        int out_row = m0 + row, rb = 0;
        bool row_ok = out_row < p.co;
        if (p.row_block) {                                   // rows = (ky, co); row_block is 32 or 64, so rb is warp-uniform
            rb = row / p.row_block;
            out_row = row - rb * p.row_block;
            row_ok = out_row < p.co;
        }
        mbar_wait(smem_u32(done_bar), 0, p.err, 304);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int t = 0; t < p.n_atoms; ++t) {
            const WnAtom& at = p.atom[t];
            const int ky = at.ky_base - rb;
            const bool ok = row_ok && (!p.row_block || (ky >= 0 && ky <= 2));
            const int tap = p.row_block ? ky * 3 + at.tap : at.tap + (p.tap_by_ts ? ts : 0);
            float* dst = p.dw + ((size_t)(ok ? tap : 0) * p.ci_total + at.ci0 + nt * p.ci_tile) * p.co_pad + out_row;
            for (int j = 0; j < at.width; j += 16) {
                uint32_t v[16];
                if (p.dbg & 4) continue;
                tc_ld16(taddr + (uint32_t)(at.col + j), v);
                tc_ld_wait();
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(dst + (size_t)(j + i) * p.co_pad, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 get_encode3() {
    static EncodeTiledFn3 fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn3>(q);
    }
    return fn;
}
// NHWC bf16 [n][h][w][c_stride] viewed as (C, W, H, N); box (cw channels, 16, box_h, 1) pixels, traversal stride `stride` in x and y
static int make_nhwc_map(CUtensorMap* tm, const void* ptr, int n, int h, int w, int c_stride, int cw, int box_h, int stride) {
    EncodeTiledFn3 enc = get_encode3();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)c_stride, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c_stride * 2, (cuuint64_t)w * c_stride * 2, (cuuint64_t)h * w * c_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)(kWnTileW * stride), (cuuint32_t)(box_h * stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    const int swz = cw * 2;
    const CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { char b[128]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled(wgrad NHWC operand) failed: %d", (int)r); return fail(b); }
    return 0;
}

static int* g_wn_err = nullptr;

}  // namespace pnnp

using namespace pnnp;

// mode 0: 3x3 stride-1 pad-1 conv  — g: NHWC [n][h][w][co_stride], x: NHWC [n][h][w][ci_stride]; dw: [9][ci_total][co_pad]
// mode 1: ConvTranspose2d(2, s2)   — g: NHWC [n][2h][2w][co_stride], x: NHWC [n][h][w][ci_stride]; dw: [4][ci_total][co_pad]
// Accumulates (+=) rows [ci_off, ci_off + ci) x columns [0, co) of every tap; co, ci multiples of 16 (ci a power-of-two multiple
// of 16 up to 64, or a multiple of 128 above... see the checks).
extern "C" int pnnp_wgrad_nhwc(int mode, const void* g, int co, int co_stride, const void* x, int ci, int ci_stride, int n, int h, int w,
                               float* dw, int ci_off, int ci_total, int co_pad, void* stream) {
    if (!g || !x || !dw) return fail("wgrad_nhwc: null pointer");
    if (mode != 0 && mode != 1) return fail("wgrad_nhwc: mode must be 0 (conv3x3) or 1 (convT2x2)");
    if (co < 16 || (co % 16) || ci < 16 || (ci % 16) || co_pad < co || ci_off + ci > ci_total) return fail("wgrad_nhwc: bad channel geometry");
    if (!((ci == 16 || ci == 32 || ci == 64) || (ci % 128 == 0))) return fail("wgrad_nhwc: ci must be 16, 32, 64 or a multiple of 128");
    if (!((co == 16 || co == 32 || co == 64) || (co % 128 == 0))) return fail("wgrad_nhwc: co must be 16, 32, 64 or a multiple of 128");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    PNNP_CUDA(cudaGetDevice(&dev));
    PNNP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!g_wn_err) { PNNP_CUDA(cudaMalloc(&g_wn_err, sizeof(int))); PNNP_CUDA(cudaMemset(g_wn_err, 0, sizeof(int))); }
    static bool attr_done = false;
    if (!attr_done) { PNNP_CUDA(cudaFuncSetAttribute(wgrad_nhwc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_done = true; }

    const int cw_g = std::min(co, 64), cw_x = std::min(ci, 64);          // channels per TMA box = swizzle span / 2
    const int pitch_a = cw_g * 2, pitch_b = cw_x * 2;
    // 3x3 layers with 32 output channels (the full-resolution layers): M = 128 would be 3/4 padding.  There the filter ROW shift
    // goes to the g side instead — one haloed g box, M blocks = (ky, co) through LBO = 16 pixel rows — and the filter COLUMN
    // shift to the x side (three boxes, N = (kx, ci)): all nine taps in ONE MMA per 16 pixels, a third of the tensor work.
    const bool rows_ky = mode == 0 && (co == 32 || co == 64) && ci <= 64 && !getenv("PNNP_WG_NO_ROWS_KY");
    const int box_h_x = (mode == 0 && !rows_ky) ? kWnTileH + 2 : kWnTileH;
    const int a_box_bytes = kWnPix * pitch_a;                              // one g block: 128 pixels
    const int b_box_bytes = kWnTileW * box_h_x * pitch_b;                  // one x block: 160 (haloed) or 128 pixels
    const int a_region = 128 / cw_g * a_box_bytes;                         // M = 128 rows always addressable (rows >= co are never stored)
    CUtensorMap tmG, tmX;
    if (int e = make_nhwc_map(&tmG, g, n, mode == 0 ? h : 2 * h, mode == 0 ? w : 2 * w, co_stride, cw_g, rows_ky ? kWnTileH + 2 : kWnTileH,
                              mode == 0 ? 1 : 2)) return e;
    if (int e = make_nhwc_map(&tmX, x, n, h, w, ci_stride, cw_x, box_h_x, 1)) return e;

    const int tiles_x = (w + kWnTileW - 1) / kWnTileW, tiles_y = (h + kWnTileH - 1) / kWnTileH;
    const int ci_tile = ci <= 64 ? ci : 128;
    const int n_tiles = ci / ci_tile, m_tiles = (co + 127) / 128;
    // tap sets: 3x3 with ci <= 32 keeps all nine taps (three dx boxes) in one CTA; otherwise one filter column per CTA;
    // the transposed conv takes one of its four taps per CTA
    const int tapsets = mode == 1 ? 4 : ((ci <= 32 || rows_ky) ? 1 : 3);
    const int combos = m_tiles * n_tiles * tapsets;
    const int tiles_total = n * tiles_x * tiles_y;
    const int splits = std::max(1, std::min(tiles_total, (2 * sms + combos - 1) / combos));
    const char* e = getenv("PNNP_WG_DBG");
    const int dbg = e ? atoi(e) : 0;

    WnParams p{};
    p.pitch_a = pitch_a; p.pitch_b = pitch_b; p.swz_a = pitch_a; p.swz_b = pitch_b; p.lbo_a = a_box_bytes;
    p.tiles_x = tiles_x; p.tiles_y = tiles_y; p.tiles_total = tiles_total; p.splits = splits; p.tapsets = tapsets;
    p.n_tiles = n_tiles; p.ci_tile = ci_tile; p.tap_by_ts = tapsets > 1 ? 1 : 0;
    p.co = co; p.ci_total = ci_total; p.co_pad = co_pad; p.dw = dw; p.dbg = dbg; p.err = g_wn_err;
    int off = 0, nb = 0, col = 0;
    if (rows_ky) {
        WnBox& gb = p.box[nb++];
        gb.is_x = 0; gb.c0 = 0; gb.mul = 1; gb.dx = 0; gb.dy = -1; gb.by_ts = 0; gb.smem_off = 0;
        gb.bytes = kWnTileW * (kWnTileH + 2) * pitch_a;                 // haloed in y: tile rows -1 .. 8
        p.lbo_a = kWnTileW * pitch_a;                                   // M blocks = filter rows: 16 pixel rows apart
        p.row_block = co;                                               // 32: blocks ky = 2,1,0,(unused); 64: two MMAs (2,1) and (0,unused)
        const int n_mma = co == 32 ? 1 : 2, blocks_per_mma = 128 / co;
        // M = 128 rows from the last MMA's first block stay inside the stage
        off = (((n_mma - 1) * blocks_per_mma * kWnTileW + (blocks_per_mma - 1) * kWnTileW + kWnPix) * pitch_a + 1023) / 1024 * 1024;
        const int xb = kWnPix * pitch_b;
        for (int dx = 0; dx < 3; ++dx) {
            WnBox& bx = p.box[nb++];
            bx.is_x = 1; bx.c0 = 0; bx.mul = 1; bx.dx = dx - 1; bx.dy = 0; bx.by_ts = 0; bx.smem_off = off + dx * xb; bx.bytes = xb;
        }
        for (int m = 0; m < n_mma; ++m) {
            WnMma& mm = p.mma[p.n_mmas++];
            mm.a_off = m * blocks_per_mma * kWnTileW * pitch_a;           // first block of MMA m starts at box row m * blocks_per_mma
            mm.b_off = off; mm.n = 3 * ci; mm.lbo_b = xb; mm.tmem_col = m * 3 * ci;
            for (int dx = 0; dx < 3; ++dx)                               // box row r of the haloed g tile <-> ky = 2 - r; tap field = kx
                p.atom[p.n_atoms++] = WnAtom{m * 3 * ci + dx * ci, ci, dx, ci_off, 2 - m * blocks_per_mma};
        }
        col = n_mma * 3 * ci; off += 3 * xb;
    }
    // g blocks of an M tile (first channel advances by 128 per M tile in the kernel)
    const int g_blocks = rows_ky ? 0 : std::min(128, co) / cw_g;
    for (int b = 0; b < g_blocks; ++b) {
        WnBox& bx = p.box[nb++];
        bx.is_x = 0; bx.c0 = b * cw_g; bx.smem_off = b * a_box_bytes; bx.bytes = a_box_bytes;
        bx.mul = mode == 0 ? 1 : 2; bx.dx = 0; bx.dy = 0; bx.by_ts = mode == 0 ? 0 : 2;      // transposed conv: rows 2y + a, cols 2x + b
    }
    if (!rows_ky) off = a_region;
    if (rows_ky) {
        // planned above
    } else if (mode == 0 && ci <= 32) {                                // all nine taps: three filter-column boxes, dy folded into N
        for (int dx = 0; dx < 3; ++dx) {
            WnBox& bx = p.box[nb++];
            bx.is_x = 1; bx.c0 = 0; bx.mul = 1; bx.dx = dx - 1; bx.dy = -1; bx.by_ts = 0; bx.smem_off = off; bx.bytes = b_box_bytes;
            WnMma& mm = p.mma[p.n_mmas++];
            mm.a_off = 0; mm.b_off = off; mm.n = 3 * ci; mm.lbo_b = kWnTileW * pitch_b; mm.tmem_col = col;
            for (int dy = 0; dy < 3; ++dy) p.atom[p.n_atoms++] = WnAtom{col + dy * ci, ci, dy * 3 + dx, ci_off, 0};
            col += 3 * ci; off += b_box_bytes;
        }
    } else if (mode == 0 && ci == 64) {                                // one filter column (ts) per CTA, dy folded into N = 192
        WnBox& bx = p.box[nb++];
        bx.is_x = 1; bx.c0 = 0; bx.mul = 1; bx.dx = 0; bx.dy = -1; bx.by_ts = 1; bx.smem_off = off; bx.bytes = b_box_bytes;
        WnMma& mm = p.mma[p.n_mmas++];
        mm.a_off = 0; mm.b_off = off; mm.n = 192; mm.lbo_b = kWnTileW * pitch_b; mm.tmem_col = 0;
        for (int dy = 0; dy < 3; ++dy) p.atom[p.n_atoms++] = WnAtom{dy * 64, 64, dy * 3, ci_off, 0};
        col = 192; off += b_box_bytes;
    } else if (mode == 0) {                                            // one filter column per CTA, 128 input channels, one MMA per dy
        for (int c = 0; c < 2; ++c) {
            WnBox& bx = p.box[nb++];
            bx.is_x = 1; bx.c0 = c * 64; bx.mul = 1; bx.dx = 0; bx.dy = -1; bx.by_ts = 1; bx.smem_off = off + c * b_box_bytes; bx.bytes = b_box_bytes;
        }
        for (int dy = 0; dy < 3; ++dy) {
            WnMma& mm = p.mma[p.n_mmas++];
            mm.a_off = 0; mm.b_off = off + dy * kWnTileW * pitch_b; mm.n = 128; mm.lbo_b = b_box_bytes; mm.tmem_col = dy * 128;
            for (int c = 0; c < 2; ++c) p.atom[p.n_atoms++] = WnAtom{dy * 128 + c * 64, 64, dy * 3, ci_off + c * 64, 0};
        }
        col = 384; off += 2 * b_box_bytes;
    } else {                                                           // transposed conv: one tap (ts) per CTA
        const int blocks = ci_tile / cw_x;
        for (int c = 0; c < blocks; ++c) {
            WnBox& bx = p.box[nb++];
            bx.is_x = 1; bx.c0 = c * cw_x; bx.mul = 1; bx.dx = 0; bx.dy = 0; bx.by_ts = 0; bx.smem_off = off + c * b_box_bytes; bx.bytes = b_box_bytes;
            p.atom[p.n_atoms++] = WnAtom{c * cw_x, cw_x, 0, ci_off + c * cw_x, 0};
        }
        WnMma& mm = p.mma[p.n_mmas++];
        mm.a_off = 0; mm.b_off = off; mm.n = ci_tile; mm.lbo_b = b_box_bytes; mm.tmem_col = 0;
        col = ci_tile; off += blocks * b_box_bytes;
    }
    p.n_boxes = nb;
    for (int b = 0; b < nb; ++b) p.stage_tx += p.box[b].bytes;
    p.stage_bytes = (off + 1023) / 1024 * 1024;
    p.stages = std::max(2, std::min(kWnStagesMax, (227 * 1024 - 2048) / p.stage_bytes));
    int tc = 32; while (tc < col) tc <<= 1;
    p.tmem_cols = tc;
    const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + (2 * kWnStagesMax + 1) * 8 + 64;
    if (smem > 227 * 1024 || tc > 512) return fail("wgrad_nhwc: shared memory / TMEM budget exceeded");
    if (variant_on("PNNP_WGRAD_V2") && !p.dbg) {
        static bool attr2_done = false;
        if (!attr2_done) { PNNP_CUDA(cudaFuncSetAttribute(wgrad_nhwc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr2_done = true; }
        PNNP_WGRAD_KLAUNCH(1);
    } else
        PNNP_WGRAD_KLAUNCH(0);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_wgrad_nhwc_pipeline_error(void) {
    int v = 0;
    if (g_wn_err) { cudaMemcpy(&v, g_wn_err, sizeof(int), cudaMemcpyDeviceToHost); if (v) cudaMemset(g_wn_err, 0, sizeof(int)); }
    return v;
}
