// Separable form of the SSIM / squared-error pass of eval_metrics.cu (the default since r02; PNNP_SSIM_V2=0: the first form).
//
// The 7x7 uniform-window sums of skimage's structural_similarity (utils/visualization.py:29-30) are separable: for each of the
// window quantities (a, b, a^2 + b^2, ab; five with a^2 and b^2 apart in the first version) a horizontal 7-sum per patch row, then a vertical 7-sum per window centre — 2 x 7 float64
// additions per quantity and centre instead of 49.  The first version (ssim_mse_kernel) spends ~440 float64-heavy instructions
// per pixel, about 1 ms of the 3.8 ms evaltest frame.
//
// The kernel body is written as PHASES separated by __syncthreads(): each phase is a function of (thread id, shared tile), so the
// CPU suite can run phase after phase over all threads of a block (tests/emul/, test infrastructure only) and compare the very
// same source with the oracle.
#pragma once
#include <cstddef>
#include <cstdint>
#ifndef PNNP_HOST_EMUL
#include <cuda_runtime.h>
#endif

namespace pnnp {

constexpr int kS2Win = 7, kS2Pad = 3;
constexpr int kS2TileX = 32, kS2TileY = 16;                     // window centres per block
constexpr int kS2PatchX = kS2TileX + 2 * kS2Pad, kS2PatchY = kS2TileY + 2 * kS2Pad;   // 38 x 22 pixels
// 128 threads: both summing passes then have exactly one item per thread and tile — 16 new patch rows x 8 column groups in the
// horizontal pass (with 256 threads half of the block idled there), 4 row groups x 32 columns in the vertical pass, whose items
// cover FOUR centre rows (ten rows of horizontal sums per quantity for four outputs instead of eight for two: 2.5 instead of 4
// 8-byte shared-memory loads per output and quantity; the pass is bound by shared-memory bytes — ~180 per output before, r02).
// PNNP_S2_THREADS=256 / PNNP_S2_VROWS=2 rebuild the earlier form (same-box comparisons).
#ifndef PNNP_S2_THREADS
#define PNNP_S2_THREADS 128
#endif
#ifndef PNNP_S2_VROWS
#define PNNP_S2_VROWS 4
#endif
constexpr int kS2Threads = PNNP_S2_THREADS;
constexpr int kS2VRows = PNNP_S2_VROWS;
static_assert(kS2VRows == 2 || kS2VRows == 4, "vertical-pass items cover two or four centre rows");
// A block walks kS2TilesPerCta vertically adjacent tiles.  The bottom six patch rows of a tile are the top six of the next one, so
// from the second tile on only the 16 NEW rows are loaded and summed horizontally (22 before: 27 % of the load, conversion,
// multiplication and horizontal-sum work); the horizontal sums live in a ring of kS2Ring rows indexed by the row's distance from
// the block's first patch row (a power of two: the slot is an AND).
constexpr int kS2Ring = 32;

struct Ssim2Args {
    const float* dn; const float* hr;
    int c, h, w;
    float scale, gain;                                          // gain applied when use_gain (IlluminanceCorrect)
    int use_gain;
};

struct Ssim2Tile {
    float a[kS2PatchY][kS2PatchX + 1], b[kS2PatchY][kS2PatchX + 1];
    // horizontal 7-sums per patch row and centre column: 28 KB.  Columns are stored permuted (ssim2_col): the four sums a thread of the
    // horizontal pass produces go out as two 16-byte stores, and with the natural order neighbouring lanes were 32 bytes apart — a
    // quarter-warp's 128-bit store then spans 256 bytes, two wavefronts instead of one (r02 capture: 35 M store bank conflicts, the
    // pass bound by shared-memory wavefronts at 62 % of the LSU pipe).  Sums 0-1 of column group g sit at [2g, 2g + 1], sums 2-3 at
    // [16 + 2g, 16 + 2g + 1]: lanes 16 bytes apart in both stores; the vertical pass walks the columns in the stored order.
    // FOUR quantities: sum a, sum b, sum (a^2 + b^2), sum ab — the SSIM map needs the two variances only as their sum, so a^2 and b^2
    // share one window sum (a fifth of the additions, shared-memory stores and loads of both passes; the first form kept five)
    double hs[4][kS2Ring][kS2TileX];
};
__device__ __forceinline__ int ssim2_col(int lx) { return ((lx >> 1) & 1) * 16 + (lx >> 2) * 2 + (lx & 1); }

// phase 1, split in two so that the kernel can keep the NEXT tile's global loads in flight while it computes the current one (r02
// capture: the pass spent 4.3 of its 14 warp cycles per issue waiting for these loads and 4.5 at the barriers behind them):
//   ssim2_fetch  the thread's (up to) four pixels of both planes into registers, zero outside the image;
//   ssim2_stage  tensor2im of those registers into the shared tile; returns this thread's share of the squared error.
constexpr int kS2Slots = (kS2PatchY * kS2PatchX + kS2Threads - 1) / kS2Threads;        // 4
struct Ssim2Regs { float d[kS2Slots], r[kS2Slots]; };
// py0: first patch row to load (0 for a block's first tile, 6 for the following ones: rows 0-5 were the previous tile's rows 16-21)
// INSIDE: the whole patch (rows py0 .. 21, all 38 columns) lies inside the image — the caller's block-uniform test — so the four
// per-pixel bounds comparisons of both halves go away (r02 capture: 16 ISETP per output, most of them these; four tiles in five of
// a 512 x 512 crop are interior).
template <bool INSIDE = false>
__device__ __forceinline__ void ssim2_fetch(int tid, const Ssim2Args& g, int plane, int x0, int y0, Ssim2Regs& v, int py0 = 0) {
    const float* d = g.dn + (size_t)plane * g.h * g.w;
    const float* r = g.hr + (size_t)plane * g.h * g.w;
#pragma unroll
    for (int k = 0; k < kS2Slots; ++k) {
        const int i = tid + k * kS2Threads + py0 * kS2PatchX;
        const int py = i / kS2PatchX, px = i - py * kS2PatchX;
        const int gx = x0 + px - kS2Pad, gy = y0 + py - kS2Pad;
        const bool in = i < kS2PatchY * kS2PatchX && (INSIDE || (gx >= 0 && gx < g.w && gy >= 0 && gy < g.h));
        v.d[k] = in ? d[(size_t)gy * g.w + gx] : 0.f;
        v.r[k] = in ? r[(size_t)gy * g.w + gx] : 0.f;
    }
}
__device__ __forceinline__ bool ssim2_inside(const Ssim2Args& g, int x0, int y0, int py0) {
    return x0 - kS2Pad >= 0 && x0 + kS2TileX + kS2Pad <= g.w && y0 - kS2Pad + py0 >= 0 && y0 + kS2TileY + kS2Pad <= g.h;
}
// se_lo / se_hi: image rows [se_lo, se_hi) whose squared error this call accounts for (a block's first tile: its own 16 centre rows
// and, like every tile, the three rows below them when another tile of the block follows — those rows are not staged again)
template <bool INSIDE = false>
__device__ __forceinline__ double ssim2_stage(int tid, const Ssim2Args& g, int x0, int y0, const Ssim2Regs& v, Ssim2Tile& t, int py0 = 0,
                                              int se_lo = -1, int se_hi = -1) {
    if (se_lo < 0) { se_lo = y0; se_hi = y0 + kS2TileY; }
    double se = 0.0;
#pragma unroll
    for (int k = 0; k < kS2Slots; ++k) {
        const int i = tid + k * kS2Threads + py0 * kS2PatchX;
        if (i >= kS2PatchY * kS2PatchX) break;
        const int py = i / kS2PatchX, px = i - py * kS2PatchX;
        const int gx = x0 + px - kS2Pad, gy = y0 + py - kS2Pad;
        float a = 0.f, b = 0.f;
        if (INSIDE || (gx >= 0 && gx < g.w && gy >= 0 && gy < g.h)) {
            float p = fminf(fmaxf(v.d[k] * g.scale, 0.f), 1.f);
            if (g.use_gain) p = g.gain * p;
            a = fminf(fmaxf(p * 255.0f, 0.f), 255.f);
            b = fminf(fmaxf(v.r[k] * 255.0f, 0.f), 255.f);
            if (px >= kS2Pad && px < kS2TileX + kS2Pad && gy >= se_lo && gy < se_hi) {
                const double e = (double)b - (double)a;
                se += e * e;
            }
        }
        t.a[py][px] = a;
        t.b[py][px] = b;
    }
    return se;
}
// both halves in one call (the CPU suite's phase-by-phase driver)
__device__ __forceinline__ double ssim2_load(int tid, const Ssim2Args& g, int plane, int x0, int y0, Ssim2Tile& t) {
    Ssim2Regs v;
    ssim2_fetch(tid, g, plane, x0, y0, v);
    return ssim2_stage(tid, g, x0, y0, v, t);
}

// Four adjacent 7-sums w_k = v[k] + ... + v[k + 6] (k = 0..3) of ten values share their terms: the core v[3..6] is common to all
// four, v[1] + v[2] to w0 and w1, v[7] + v[8] to w2 and w3 — 13 additions instead of 24, only additions (no running-sum subtraction,
// hence no cancellation), each w_k still the sum of its own seven values in a fixed association.
__device__ __forceinline__ void ssim2_sums4(const double (&v)[10], double (&w)[4]) {
    const double core = (v[3] + v[4]) + (v[5] + v[6]);
    const double p12 = v[1] + v[2], p78 = v[7] + v[8];
    w[0] = core + (v[0] + p12);
    w[1] = core + (p12 + v[7]);
    w[2] = core + (v[2] + p78);
    w[3] = core + (p78 + v[9]);
}

// the four sums of column group g = lx / 4 of one patch row as two 16-byte stores in the permuted column order
__device__ __forceinline__ void ssim2_store4(double* row, int g, const double (&w)[4]) {
    reinterpret_cast<double2*>(row)[g] = make_double2(w[0], w[1]);
    reinterpret_cast<double2*>(row)[8 + g] = make_double2(w[2], w[3]);
}

// phase 2: horizontal 7-sums; one item = (patch row, group of FOUR adjacent centre columns): ten pixels loaded, converted and multiplied
// once (the first form loaded, converted and multiplied every pixel seven times: 70 float64 operations per output, 29 now)
// py0: first patch row to sum (see ssim2_fetch); ring0: ring slot of patch row 0 of this tile
__device__ __forceinline__ void ssim2_hsum(int tid, Ssim2Tile& t, int py0 = 0, int ring0 = 0) {
    constexpr int kGroups = kS2TileX / 4;
    for (int i = tid + py0 * kGroups; i < kS2PatchY * kGroups; i += kS2Threads) {
        const int py = i / kGroups, lx = (i - py * kGroups) * 4, slot = (ring0 + py) & (kS2Ring - 1);
        double a[10], b[10], pp[10], ab[10], w[4];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            a[k] = t.a[py][lx + k]; b[k] = t.b[py][lx + k];
            pp[k] = fma(a[k], a[k], b[k] * b[k]); ab[k] = a[k] * b[k];      // a^2, b^2 are exact (24-bit significands): one rounding
        }
        ssim2_sums4(a, w);
        ssim2_store4(t.hs[0][slot], lx >> 2, w);
        ssim2_sums4(b, w);
        ssim2_store4(t.hs[1][slot], lx >> 2, w);
        ssim2_sums4(pp, w);
        ssim2_store4(t.hs[2][slot], lx >> 2, w);
        ssim2_sums4(ab, w);
        ssim2_store4(t.hs[3][slot], lx >> 2, w);
    }
}

// 1 / d for the map's denominator (d >= C1 C2 > 0) without the division subroutine: the hardware's 2^-23 seed and two Newton steps
// (2^-46, then the float64 rounding level); a handful of FMAs against ~15 instructions and a slow-path test per pixel
__device__ __forceinline__ double ssim2_rcp(double d) {
#ifdef PNNP_HOST_EMUL
    return 1.0 / d;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(r, fma(-d, r, 1.0), r);
    r = fma(r, fma(-d, r, 1.0), r);
    return r;
#endif
}

// phase 3: vertical 7-sums + the SSIM map value; one item = (centre column, kS2VRows adjacent centre rows) = one per thread (an
// earlier form's items of four rows in a 256-thread block kept half of the block idle during the pass's heaviest phase: barrier
// stalls 4.5 warp cycles per issue in the r02 capture); ten (eight) rows of horizontal sums per quantity, their common rows summed
// once; returns this thread's share of the map's sum
__device__ __forceinline__ double ssim2_vsum(int tid, const Ssim2Args& g, int x0, int y0, const Ssim2Tile& t, int ring0 = 0) {
    const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
    const double inv_np = 1.0 / 49.0, cov_norm = 49.0 / 48.0;
    double ssum = 0.0;
    for (int i = tid; i < (kS2TileY / kS2VRows) * kS2TileX; i += kS2Threads) {
        // consecutive lanes take columns that are consecutive IN THE STORED ORDER (a half-warp's 64-bit load is one 128-byte run);
        // lx = ssim2_col^-1(col) is the centre column those sums belong to
        const int lg = i / kS2TileX, col = i - lg * kS2TileX, ly0 = lg * kS2VRows;
        const int lx = ((col & 15) >> 1) * 4 + (col >> 4) * 2 + (col & 1), cx = x0 + lx;
        if (cx < kS2Pad || cx >= g.w - kS2Pad) continue;
        double s[4][kS2VRows];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if constexpr (kS2VRows == 4) {
                double v[10], w[4];
#pragma unroll
                for (int k = 0; k < 10; ++k) v[k] = t.hs[q][(ring0 + ly0 + k) & (kS2Ring - 1)][col];
                ssim2_sums4(v, w);                          // the horizontal pass's sharing of terms, down a column
#pragma unroll
                for (int k = 0; k < 4; ++k) s[q][k] = w[k];
            } else {
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = t.hs[q][(ring0 + ly0 + k) & (kS2Ring - 1)][col];
                const double core = ((v[1] + v[2]) + (v[3] + v[4])) + (v[5] + v[6]);
                s[q][0] = core + v[0];
                s[q][1] = core + v[7];
            }
        }
#pragma unroll
        for (int k = 0; k < kS2VRows; ++k) {
            const int cy = y0 + ly0 + k;
            if (cy < kS2Pad || cy >= g.h - kS2Pad) continue;
            const double ux = s[0][k] * inv_np, uy = s[1][k] * inv_np;
            const double uxuy = ux * uy, m2 = fma(ux, ux, uy * uy);
            const double vxvy = cov_norm * (s[2][k] * inv_np - m2);            // var a + var b (sample variances)
            const double vxy = cov_norm * (s[3][k] * inv_np - uxuy);
            ssum += ((2 * uxuy + C1) * (2 * vxy + C2)) * ssim2_rcp((m2 + C1) * (vxvy + C2));
        }
    }
    return ssum;
}

}  // namespace pnnp
