// Kernels of the eval epilogue (launcher and the C ABI: eval_metrics.cu).  In a header so that the CPU suite can compile this very
// source for the host and run whole CTAs of it on the lock-step fibre emulator (tests/emul/simt_host.h; test infrastructure only).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>
#include "ssim_core.cuh"

#ifndef PNNP_SMEM
#ifdef PNNP_HOST_EMUL
#define PNNP_SMEM static                 // one CTA at a time on the host: a static array is the CTA's shared memory
#else
#define PNNP_SMEM __shared__
#endif
#endif

namespace pnnp {

constexpr int kSsimWin = 7, kSsimPad = 3;
constexpr int kTileX = 32, kTileY = 32;

__device__ __forceinline__ double block_reduce_sum(double v, double* s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? s_red[threadIdx.x] : 0.0;
    if (warp == 0) for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;     // valid in thread 0
}

// pass 1: illuminance dots per frame
__global__ void __launch_bounds__(256) illum_dots_kernel(const float* __restrict__ dn, const float* __restrict__ hr, size_t per_frame,
                                                         float scale, double* sums, int stride) {
    PNNP_SMEM double s_red[8];
    const int frame = blockIdx.y;
    const float* d = dn + (size_t)frame * per_frame;
    const float* t = hr + (size_t)frame * per_frame;
    double num = 0.0, den = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_frame; i += (size_t)gridDim.x * blockDim.x) {
        const float p = fminf(fmaxf(d[i] * scale, 0.f), 1.f);
        const float s = t[i];
        if (s != 1.0f) { num += (double)p * (double)s; den += (double)p * (double)p; }
    }
    num = block_reduce_sum(num, s_red);
    den = block_reduce_sum(den, s_red);
    if (threadIdx.x == 0) { atomicAdd(&sums[frame * stride + 0], num); atomicAdd(&sums[frame * stride + 1], den); }
}

// pass 2: squared error + SSIM map sums.  One block = one 32x32 tile of one (frame, channel) plane.
__global__ void __launch_bounds__(256) ssim_mse_kernel(const float* __restrict__ dn, const float* __restrict__ hr, int c, int h, int w,
                                                       float scale, int use_gain, double* sums, int stride) {
    PNNP_SMEM float sx[kTileY + 2 * kSsimPad][kTileX + 2 * kSsimPad + 1];
    PNNP_SMEM float sy[kTileY + 2 * kSsimPad][kTileX + 2 * kSsimPad + 1];
    PNNP_SMEM double s_red[8];
    const int plane_id = blockIdx.z;                 // frame * c + channel
    const int frame = plane_id / c, ch = plane_id - frame * c;
    const float* d = dn + (size_t)plane_id * h * w;
    const float* t = hr + (size_t)plane_id * h * w;
    float gain = 1.f;
    if (use_gain) gain = (float)sums[frame * stride + 0] / (float)sums[frame * stride + 1];   // num / den in float32 like torch
    const int x0 = blockIdx.x * kTileX, y0 = blockIdx.y * kTileY;
    // tile (x0..x0+32, y0..y0+32) of window CENTRES; patch covers centres +/- 3
    double se = 0.0;
    for (int i = threadIdx.x; i < (kTileY + 6) * (kTileX + 6); i += blockDim.x) {
        const int py = i / (kTileX + 6), px = i - py * (kTileX + 6);
        const int gx = x0 + px - kSsimPad, gy = y0 + py - kSsimPad;
        float a = 0.f, b = 0.f;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
            float p = fminf(fmaxf(d[(size_t)gy * w + gx] * scale, 0.f), 1.f);
            if (use_gain) p = gain * p;
            a = fminf(fmaxf(p * 255.0f, 0.f), 255.f);                       // tensor2im(estimate)
            b = fminf(fmaxf(t[(size_t)gy * w + gx] * 255.0f, 0.f), 255.f);  // tensor2im(target)
            // each pixel belongs to exactly one tile's interior: count its squared error there
            if (px >= kSsimPad && px < kTileX + kSsimPad && py >= kSsimPad && py < kTileY + kSsimPad) {
                const double e = (double)b - (double)a;
                se += e * e;
            }
        }
        sx[py][px] = a;
        sy[py][px] = b;
    }
    __syncthreads();
    const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
    const double inv_np = 1.0 / 49.0, cov_norm = 49.0 / 48.0;
    double ssum = 0.0;
    for (int i = threadIdx.x; i < kTileX * kTileY; i += blockDim.x) {
        const int ly = i / kTileX, lx = i - ly * kTileX;
        const int cx = x0 + lx, cy = y0 + ly;
        if (cx < kSsimPad || cx >= w - kSsimPad || cy < kSsimPad || cy >= h - kSsimPad) continue;
        double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
#pragma unroll
        for (int dy = 0; dy < kSsimWin; ++dy)
#pragma unroll
            for (int dx = 0; dx < kSsimWin; ++dx) {
                const double a = sx[ly + dy][lx + dx], b = sy[ly + dy][lx + dx];
                sa += a; sb += b; saa += a * a; sbb += b * b; sab += a * b;
            }
        const double ux = sa * inv_np, uy = sb * inv_np;
        const double vx = cov_norm * (saa * inv_np - ux * ux), vy = cov_norm * (sbb * inv_np - uy * uy);
        const double vxy = cov_norm * (sab * inv_np - ux * uy);
        ssum += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
    se = block_reduce_sum(se, s_red);
    ssum = block_reduce_sum(ssum, s_red);
    if (threadIdx.x == 0) { atomicAdd(&sums[frame * stride + 2], se); atomicAdd(&sums[frame * stride + 3 + ch], ssum); }
}

// pass 2, separable form (ssim_core.cuh): one block = kS2TilesPerCta vertically adjacent 32 x 16 tiles of window centres of one
// (frame, channel) plane.  The two partial sums stay in registers across the block's tiles and are reduced and added to the frame's
// accumulators ONCE per block: with one tile per block every block ended in two float64 atomics on the same two addresses per frame
// (65 536 same-address atomics per 16 crops serialise in L2 — the whole kernel's 291 us in r02, whatever the arithmetic cost).
constexpr int kS2TilesPerCta = 8;
__global__ void __launch_bounds__(kS2Threads, kS2Threads == 128 ? 5 : 4) ssim_mse_v2_kernel(Ssim2Args g, double* sums, int stride) {
    extern __shared__ __align__(16) uint8_t s2_raw[];
    Ssim2Tile& t = *reinterpret_cast<Ssim2Tile*>(s2_raw);
    PNNP_SMEM double s_red[8];
    const int plane_id = blockIdx.z, frame = plane_id / g.c, ch = plane_id - frame * g.c;
    if (g.use_gain) g.gain = (float)sums[frame * stride + 0] / (float)sums[frame * stride + 1];   // num / den in float32 like torch
    const int x0 = blockIdx.x * kS2TileX;
    double se = 0.0, ssum = 0.0;
    Ssim2Regs nxt;                                      // the next tile's pixels: loaded while the current tile is summed
    const int y_first = blockIdx.y * kS2TilesPerCta * kS2TileY;
    if (ssim2_inside(g, x0, y_first, 0)) ssim2_fetch<true>(threadIdx.x, g, plane_id, x0, y_first, nxt);
    else ssim2_fetch(threadIdx.x, g, plane_id, x0, y_first, nxt);
    for (int k = 0; k < kS2TilesPerCta; ++k) {
        const int y0 = y_first + k * kS2TileY;
        if (y0 >= g.h) break;
        // from the second tile on only the 16 new patch rows are loaded, converted and summed horizontally: rows 0-5 are the previous
        // tile's rows 16-21, still in the ring of horizontal sums
        const int py0 = k == 0 ? 0 : 2 * kS2Pad, ring0 = (k * kS2TileY) & (kS2Ring - 1);
        const bool more = k + 1 < kS2TilesPerCta && y0 + kS2TileY < g.h;
        // squared error: every image row of the block's range exactly once, by the tile that stages it (staged rows: y0 - 3 + [py0, 22))
        const int se_lo = max(y_first, y0 - kS2Pad + py0), se_hi = more ? y0 + kS2TileY + kS2Pad : y0 + kS2TileY;
        if (ssim2_inside(g, x0, y0, py0)) se += ssim2_stage<true>(threadIdx.x, g, x0, y0, nxt, t, py0, se_lo, se_hi);
        else se += ssim2_stage(threadIdx.x, g, x0, y0, nxt, t, py0, se_lo, se_hi);
        if (more) {
            if (ssim2_inside(g, x0, y0 + kS2TileY, 2 * kS2Pad)) ssim2_fetch<true>(threadIdx.x, g, plane_id, x0, y0 + kS2TileY, nxt, 2 * kS2Pad);
            else ssim2_fetch(threadIdx.x, g, plane_id, x0, y0 + kS2TileY, nxt, 2 * kS2Pad);
        }
        __syncthreads();
        ssim2_hsum(threadIdx.x, t, py0, ring0);
        __syncthreads();
        ssum += ssim2_vsum(threadIdx.x, g, x0, y0, t, ring0);
        __syncthreads();                            // the tile buffers are reused by the next tile
    }
    se = block_reduce_sum(se, s_red);
    ssum = block_reduce_sum(ssum, s_red);
    if (threadIdx.x == 0) { atomicAdd(&sums[frame * stride + 2], se); atomicAdd(&sums[frame * stride + 3 + ch], ssum); }
}

}  // namespace pnnp
