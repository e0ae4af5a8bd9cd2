// Device code of pack.cu (Bayer pack + normalisation, its inverse, the dark-shading variant), kept in a header so that the CPU suite
// can compile these very kernels for the host and run them thread by thread (tests/emul/, test infrastructure only).  Host-side
// launch code stays in pack.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#ifndef PNNP_HOST_EMUL
#include <cuda_runtime.h>
#endif
#include "pack_core.cuh"

namespace pnnp {

struct PackArgs {
    const void* raw; float* out; int n, H, W; double wp; double black[4]; int norm, clip;
};


template <typename T> struct Load8;
template <> struct Load8<uint16_t> {
    static __device__ __forceinline__ void ld(const uint16_t* p, float (&v)[8]) {
        const uint4 q = __ldcs(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = (float)(w[i] & 0xFFFFu); v[2 * i + 1] = (float)(w[i] >> 16); }
    }
};
template <> struct Load8<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(p));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) pack_norm_kernel(const PackArgs a) {
    const T* raw = static_cast<const T*>(a.raw);
    const int h = a.H / 2, w = a.W / 2;
    const size_t plane = (size_t)h * w;
    if (VEC) {
        const int w4 = w / 4;                                  // float4 groups per packed row
        const size_t total = (size_t)a.n * h * w4;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const int x4 = (int)(i % w4);
            const size_t t = i / w4;
            const int y = (int)(t % h);
            const int f = (int)(t / h);
            const T* r0 = raw + ((size_t)f * a.H + 2 * y) * a.W + 8 * x4;
            float e[8], o[8];
            Load8<T>::ld(r0, e);
            Load8<T>::ld(r0 + a.W, o);
            float4 R, G1, B, G2;
            R.x = norm_one(e[0], a.black[0], a.wp, a.norm, a.clip); R.y = norm_one(e[2], a.black[0], a.wp, a.norm, a.clip);
            R.z = norm_one(e[4], a.black[0], a.wp, a.norm, a.clip); R.w = norm_one(e[6], a.black[0], a.wp, a.norm, a.clip);
            G1.x = norm_one(e[1], a.black[1], a.wp, a.norm, a.clip); G1.y = norm_one(e[3], a.black[1], a.wp, a.norm, a.clip);
            G1.z = norm_one(e[5], a.black[1], a.wp, a.norm, a.clip); G1.w = norm_one(e[7], a.black[1], a.wp, a.norm, a.clip);
            B.x = norm_one(o[1], a.black[2], a.wp, a.norm, a.clip); B.y = norm_one(o[3], a.black[2], a.wp, a.norm, a.clip);
            B.z = norm_one(o[5], a.black[2], a.wp, a.norm, a.clip); B.w = norm_one(o[7], a.black[2], a.wp, a.norm, a.clip);
            G2.x = norm_one(o[0], a.black[3], a.wp, a.norm, a.clip); G2.y = norm_one(o[2], a.black[3], a.wp, a.norm, a.clip);
            G2.z = norm_one(o[4], a.black[3], a.wp, a.norm, a.clip); G2.w = norm_one(o[6], a.black[3], a.wp, a.norm, a.clip);
            float* ob = a.out + (size_t)f * 4 * plane + (size_t)y * w + 4 * x4;
            __stcs(reinterpret_cast<float4*>(ob), R);
            __stcs(reinterpret_cast<float4*>(ob + plane), G1);
            __stcs(reinterpret_cast<float4*>(ob + 2 * plane), B);
            __stcs(reinterpret_cast<float4*>(ob + 3 * plane), G2);
        }
    } else {
        const size_t total = (size_t)a.n * 4 * plane;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const int x = (int)(i % w);
            size_t t = i / w;
            const int y = (int)(t % h); t /= h;
            const int c = (int)(t % 4);
            const int f = (int)(t / 4);
            const int dy = (c >= 2), dx = (c == 1 || c == 2);
            const float v = (float)raw[((size_t)f * a.H + 2 * y + dy) * a.W + 2 * x + dx];
            a.out[i] = norm_one(v, a.black[c], a.wp, a.norm, a.clip);
        }
    }
}


template <bool VEC>
__global__ void __launch_bounds__(256) unpack_quant_kernel(const float* packed, uint16_t* raw, int n, int h, int w,
                                                           float span, float bl) {
    const size_t plane = (size_t)h * w;
    const int W = 2 * w;
    if (VEC) {
        const int w4 = w / 4;
        const size_t total = (size_t)n * h * w4;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const int x4 = (int)(i % w4);
            const size_t t = i / w4;
            const int y = (int)(t % h);
            const int f = (int)(t / h);
            const float* pb = packed + (size_t)f * 4 * plane + (size_t)y * w + 4 * x4;
            const float4 R = __ldcs(reinterpret_cast<const float4*>(pb));
            const float4 G1 = __ldcs(reinterpret_cast<const float4*>(pb + plane));
            const float4 B = __ldcs(reinterpret_cast<const float4*>(pb + 2 * plane));
            const float4 G2 = __ldcs(reinterpret_cast<const float4*>(pb + 3 * plane));
            uint4 e, o;
            e.x = quant_one(R.x, span, bl) | (quant_one(G1.x, span, bl) << 16);
            e.y = quant_one(R.y, span, bl) | (quant_one(G1.y, span, bl) << 16);
            e.z = quant_one(R.z, span, bl) | (quant_one(G1.z, span, bl) << 16);
            e.w = quant_one(R.w, span, bl) | (quant_one(G1.w, span, bl) << 16);
            o.x = quant_one(G2.x, span, bl) | (quant_one(B.x, span, bl) << 16);
            o.y = quant_one(G2.y, span, bl) | (quant_one(B.y, span, bl) << 16);
            o.z = quant_one(G2.z, span, bl) | (quant_one(B.z, span, bl) << 16);
            o.w = quant_one(G2.w, span, bl) | (quant_one(B.w, span, bl) << 16);
            uint16_t* r0 = raw + ((size_t)f * 2 * h + 2 * y) * W + 8 * x4;
            __stcs(reinterpret_cast<uint4*>(r0), e);
            __stcs(reinterpret_cast<uint4*>(r0 + W), o);
        }
    } else {
        const size_t total = (size_t)n * 4 * plane;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const int x = (int)(i % w);
            size_t t = i / w;
            const int y = (int)(t % h); t /= h;
            const int c = (int)(t % 4);
            const int f = (int)(t / 4);
            const int dy = (c >= 2), dx = (c == 1 || c == 2);
            raw[((size_t)f * 2 * h + 2 * y + dy) * W + 2 * x + dx] = (uint16_t)quant_one(packed[i], span, bl);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Dark-shading correction fused into the pack (the real-data side of the same boundary:
// data_process/real_datasets.py:360-372 feeding raw2bayer): per sensor sample
//     v = raw - darkshading[y][x]  [+ mean(darkshading) if 'd' in noise_code]  [+ randn * biassig ('darkshading2', train)]
// in the dark map's precision — float32 arithmetic for a float32 map, float64 for a float64 map (what NumPy's
// promotion makes of `uint16 array - map`; `ds_k * iso + ds_b + BLE` is float64 when BLE is an np.float64) — then the
// cast to float32 and the normalisation of raw2bayer.  One element per thread of the packed output.
// ------------------------------------------------------------------------------------------
template <typename D>
__global__ void __launch_bounds__(256) pack_norm_dark_kernel(const uint16_t* __restrict__ raw, const D* __restrict__ dark, float* __restrict__ out,
                                                             int n, int H, int W, double wp, double b0, double b1, double b2, double b3,
                                                             int norm, int clip, D add_mean, int use_mean, D add_bias, int use_bias) {
    const int h = H / 2, w = W / 2;
    const size_t plane = (size_t)h * w, total = (size_t)n * 4 * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        size_t t = i / w;
        const int y = (int)(t % h); t /= h;
        const int c = (int)(t % 4);
        const int f = (int)(t / 4);
        const int dy = (c >= 2), dx = (c == 1 || c == 2);
        const size_t pix = (size_t)(2 * y + dy) * W + 2 * x + dx;
        float v32;
        if (sizeof(D) == 8) {
            double v = __dsub_rn((double)raw[(size_t)f * H * W + pix], (double)dark[pix]);
            if (use_mean) v = __dadd_rn(v, (double)add_mean);
            if (use_bias) v = __dadd_rn(v, (double)add_bias);
            v32 = (float)v;
        } else {
            float v = __fsub_rn((float)raw[(size_t)f * H * W + pix], (float)dark[pix]);
            if (use_mean) v = __fadd_rn(v, (float)add_mean);
            if (use_bias) v = __fadd_rn(v, (float)add_bias);
            v32 = v;
        }
        const double black = c == 0 ? b0 : (c == 1 ? b1 : (c == 2 ? b2 : b3));
        out[i] = norm_one(v32, black, wp, norm, clip);
    }
}

}  // namespace pnnp
