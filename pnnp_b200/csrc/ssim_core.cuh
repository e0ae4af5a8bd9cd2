// Separable form of the SSIM / squared-error pass of eval_metrics.cu (OPT-IN, PNNP_SSIM_V2=1, until measured on a B200).
//
// The 7x7 uniform-window sums of skimage's structural_similarity (utils/visualization.py:29-30) are separable: for each of the
// five quantities (a, b, a^2, b^2, ab) a horizontal 7-sum per patch row, then a vertical 7-sum per window centre — 2 x 7 float64
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
constexpr int kS2Threads = 256;

struct Ssim2Args {
    const float* dn; const float* hr;
    int c, h, w;
    float scale, gain;                                          // gain applied when use_gain (IlluminanceCorrect)
    int use_gain;
};

struct Ssim2Tile {
    float a[kS2PatchY][kS2PatchX + 1], b[kS2PatchY][kS2PatchX + 1];
    double hs[5][kS2PatchY][kS2TileX];                         // horizontal 7-sums per patch row and centre column: 28 KB
};

// phase 1: load + tensor2im of the (22 x 38) patch of plane `plane`; returns this thread's share of the squared error
__device__ __forceinline__ double ssim2_load(int tid, const Ssim2Args& g, int plane, int x0, int y0, Ssim2Tile& t) {
    const float* d = g.dn + (size_t)plane * g.h * g.w;
    const float* r = g.hr + (size_t)plane * g.h * g.w;
    double se = 0.0;
    for (int i = tid; i < kS2PatchY * kS2PatchX; i += kS2Threads) {
        const int py = i / kS2PatchX, px = i - py * kS2PatchX;
        const int gx = x0 + px - kS2Pad, gy = y0 + py - kS2Pad;
        float a = 0.f, b = 0.f;
        if (gx >= 0 && gx < g.w && gy >= 0 && gy < g.h) {
            float p = fminf(fmaxf(d[(size_t)gy * g.w + gx] * g.scale, 0.f), 1.f);
            if (g.use_gain) p = g.gain * p;
            a = fminf(fmaxf(p * 255.0f, 0.f), 255.f);
            b = fminf(fmaxf(r[(size_t)gy * g.w + gx] * 255.0f, 0.f), 255.f);
            if (px >= kS2Pad && px < kS2TileX + kS2Pad && py >= kS2Pad && py < kS2TileY + kS2Pad) {
                const double e = (double)b - (double)a;
                se += e * e;
            }
        }
        t.a[py][px] = a;
        t.b[py][px] = b;
    }
    return se;
}

// phase 2: horizontal 7-sums, one (patch row, centre column) per item
__device__ __forceinline__ void ssim2_hsum(int tid, Ssim2Tile& t) {
    for (int i = tid; i < kS2PatchY * kS2TileX; i += kS2Threads) {
        const int py = i / kS2TileX, lx = i - py * kS2TileX;
        double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
#pragma unroll
        for (int dx = 0; dx < kS2Win; ++dx) {
            const double a = t.a[py][lx + dx], b = t.b[py][lx + dx];
            sa += a; sb += b; saa += a * a; sbb += b * b; sab += a * b;
        }
        t.hs[0][py][lx] = sa; t.hs[1][py][lx] = sb; t.hs[2][py][lx] = saa; t.hs[3][py][lx] = sbb; t.hs[4][py][lx] = sab;
    }
}

// phase 3: vertical 7-sums + the SSIM map value per valid window centre; returns this thread's share of the map's sum
__device__ __forceinline__ double ssim2_vsum(int tid, const Ssim2Args& g, int x0, int y0, const Ssim2Tile& t) {
    const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
    const double inv_np = 1.0 / 49.0, cov_norm = 49.0 / 48.0;
    double ssum = 0.0;
    for (int i = tid; i < kS2TileY * kS2TileX; i += kS2Threads) {
        const int ly = i / kS2TileX, lx = i - ly * kS2TileX;
        const int cx = x0 + lx, cy = y0 + ly;
        if (cx < kS2Pad || cx >= g.w - kS2Pad || cy < kS2Pad || cy >= g.h - kS2Pad) continue;
        double s[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            double v = 0.0;
#pragma unroll
            for (int dy = 0; dy < kS2Win; ++dy) v += t.hs[q][ly + dy][lx];
            s[q] = v;
        }
        const double ux = s[0] * inv_np, uy = s[1] * inv_np;
        const double vx = cov_norm * (s[2] * inv_np - ux * ux), vy = cov_norm * (s[3] * inv_np - uy * uy);
        const double vxy = cov_norm * (s[4] * inv_np - ux * uy);
        ssum += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
    return ssum;
}

}  // namespace pnnp
