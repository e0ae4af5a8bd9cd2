// Device code of crop_aug.cu (crop + augmentation gather, overlapped eval tiling and its inverse, white-balance gains), kept in a
// header so that the CPU suite can compile these very kernels for the host and run them thread by thread (tests/emul/, test
// infrastructure only: grid-stride kernels without shared memory or warp collectives execute exactly when their threads run one
// after the other).  Host-side launch code stays in crop_aug.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#ifndef PNNP_HOST_EMUL
#include <cuda_runtime.h>
#endif

namespace pnnp {
constexpr int kMaxCrops = 64;
struct CropArgs {
    const float* frame; float* out;
    int c, h, w, patch, n;
    int hs[kMaxCrops], ws[kMaxCrops], mode[kMaxCrops];
};

// source coordinate inside the crop for output (i, j) under numpy.rot90(k) followed by an optional W flip
__device__ __forceinline__ void src_coord(int i, int j, int p, int mode, int& si, int& sj) {
    if (mode >> 2) j = p - 1 - j;                 // data[..., ::-1] is applied AFTER the rotation
    switch (mode & 3) {
        case 0: si = i; sj = j; break;
        case 1: si = j; sj = p - 1 - i; break;     // rot90(k=1): out[i, j] = in[j, p-1-i]
        case 2: si = p - 1 - i; sj = p - 1 - j; break;
        default: si = p - 1 - j; sj = i; break;    // k = 3: out[i, j] = in[p-1-j, i]
    }
}

__global__ void __launch_bounds__(256) crop_aug_kernel(const CropArgs a) {
    const int p = a.patch, p4 = p / 4;
    const size_t total = (size_t)a.n * a.c * p * p4;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int j4 = (int)(t % p4);
        size_t r = t / p4;
        const int i = (int)(r % p); r /= p;
        const int ch = (int)(r % a.c);
        const int k = (int)(r / a.c);
        const float* plane = a.frame + (size_t)ch * a.h * a.w;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int si, sj;
            src_coord(i, 4 * j4 + e, p, a.mode[k], si, sj);
            v[e] = __ldg(plane + (size_t)(a.hs[k] + si) * a.w + (a.ws[k] + sj));
        }
        *reinterpret_cast<float4*>(a.out + (((size_t)k * a.c + ch) * p + i) * p + 4 * j4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// ------------------------------------------------------------------------------------------
// Overlapped tiling of a frame for tile-wise inference and its inverse
// (SynBase_Dataset.eval_crop / eval_merge, data_process/syn_datasets.py:109-159; caller trainer_SID.py:345-360).
// d = base/2, l = patch - base, nh = h/l + 1, nw = w/l + 1.  The frame is reflect-padded by d; tile (i, j) starts at
// (i*l, j*l) of the padded frame, the last row / column of tiles at (H_pad - patch) / (W_pad - patch).  The merge keeps the
// interior l x l of every tile; where regions overlap the reference's later writes win (right column, bottom row, corner).
// ------------------------------------------------------------------------------------------
struct TileGeom { int c, h, w, patch, d, l, nh, nw; };
__device__ __forceinline__ int reflect_idx(int t, int n) { t = t < 0 ? -t : t; return t >= n ? 2 * (n - 1) - t : t; }

__global__ void __launch_bounds__(256) eval_crop_kernel(const float* __restrict__ frame, float* __restrict__ tiles, const TileGeom g) {
    const int p = g.patch, p4 = p / 4;
    const size_t total = (size_t)g.nh * g.nw * g.c * p * p4;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int x4 = (int)(t % p4);
        size_t r = t / p4;
        const int y = (int)(r % p); r /= p;
        const int ch = (int)(r % g.c); r /= g.c;
        const int j = (int)(r % g.nw), i = (int)(r / g.nw);
        const int oy = i < g.nh - 1 ? i * g.l : g.h + 2 * g.d - p;       // tile origin in the padded frame
        const int ox = j < g.nw - 1 ? j * g.l : g.w + 2 * g.d - p;
        const float* row = frame + ((size_t)ch * g.h + reflect_idx(oy + y - g.d, g.h)) * g.w;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = __ldg(row + reflect_idx(ox + 4 * x4 + e - g.d, g.w));
        *reinterpret_cast<float4*>(tiles + ((((size_t)i * g.nw + j) * g.c + ch) * p + y) * p + 4 * x4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__global__ void __launch_bounds__(256) eval_merge_kernel(const float* __restrict__ tiles, float* __restrict__ frame, const TileGeom g) {
    const int p = g.patch;
    const size_t total = (size_t)g.c * g.h * g.w;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % g.w);
        size_t r = t / g.w;
        const int y = (int)(r % g.h);
        const int ch = (int)(r / g.h);
        int i, ty, j, tx;
        if (y >= g.h - g.l) { i = g.nh - 1; ty = y - (g.h - g.l) + g.d; } else { i = y / g.l; ty = y - i * g.l + g.d; }
        if (x >= g.w - g.l) { j = g.nw - 1; tx = x - (g.w - g.l) + g.d; } else { j = x / g.l; tx = x - j * g.l + g.d; }
        frame[t] = __ldg(tiles + ((((size_t)i * g.nw + j) * g.c + ch) * p + ty) * p + tx);
    }
}

// ------------------------------------------------------------------------------------------
// White-balance jitter of Raw_Dataset.__getitem__ (data_process/syn_datasets.py:313-319), in place on n x c x h x w crops:
//   hr_crops *= rgb_gain                  float32, every plane
//   hr_crops[:, ch] = hr_crops[:, ch] * g  for ch = 0 (red) and 2 (blue), g = wb[ch] / gain:
//     kind 1: g is float32 -> float32 product;  kind 2: g is float64 (np.float64 white balance, NEP 50) -> the product is
//     formed in float64 and rounded to float32 once on assignment.
// Explicit _rn intrinsics: no FMA contraction, two separately rounded float32 products like NumPy's.
struct GainArgs { float common; int kind[8]; float g32[8]; double g64[8]; };

__global__ void __launch_bounds__(256) wb_gains_kernel(float* __restrict__ data, size_t plane4, int c, size_t total4, const GainArgs a) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)((t / plane4) % (size_t)c);
        float4 v = reinterpret_cast<float4*>(data)[t];
        float* f = reinterpret_cast<float*>(&v);
        const int kind = a.kind[ch];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x = __fmul_rn(f[e], a.common);
            if (kind == 1) x = __fmul_rn(x, a.g32[ch]);
            else if (kind == 2) x = __double2float_rn(__dmul_rn((double)x, a.g64[ch]));
            f[e] = x;
        }
        reinterpret_cast<float4*>(data)[t] = v;
    }
}
}  // namespace pnnp
