// The small layout kernels next to the tcgen05 conv kernel (conv_tc.cu): network-input conversion NCHW fp32 -> NHWC16 bf16 (one- and
// four-pixel forms) and 2x2 max-pool on NHWC bf16.  Kept in a header so that the CPU suite can compile these very kernels for the
// host and run them thread by thread (tests/emul/, test infrastructure only).
#pragma once
#include <cstddef>
#include <cstdint>
#ifndef PNNP_HOST_EMUL
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#endif

namespace pnnp {

// NCHW fp32 (c <= 16 channels) -> NHWC bf16 with the channel dim zero-padded to 16 (network input)
__global__ void nchw_f32_to_nhwc16_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int n, int c,
                                               int h, int w, float scale) {
    const size_t plane = (size_t)h * w, total = (size_t)n * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t img = i / plane, pix = i - img * plane;
        uint32_t pk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float a = (2 * k < c) ? in[(img * c + 2 * k) * plane + pix] * scale : 0.f;
            const float b = (2 * k + 1 < c) ? in[(img * c + 2 * k + 1) * plane + pix] * scale : 0.f;
            const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
            pk[k] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        uint4* o = reinterpret_cast<uint4*>(out + i * 16);
        o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
}

// NCHW fp32 -> NHWC fp32 zero-padded to 16 channels (input of the fp32-storage / tf32 variant of the network)
__global__ void nchw_f32_to_nhwc16_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int c, int h, int w) {
    const size_t plane = (size_t)h * w, total = (size_t)n * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t img = i / plane, pix = i - img * plane;
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = k < c ? in[(img * c + k) * plane + pix] : 0.f;
        float4* o = reinterpret_cast<float4*>(out + i * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
}

// Same conversion, four pixels per thread (h * w % 4 == 0): one 128-bit load per plane (all issued before the first use) and a
// contiguous 128-byte store per thread.  The default since r02 (55 -> 50 us per Sony frame; PNNP_IN_V2=0 for the first form): the one-pixel kernel above takes 53 us for a Sony
// frame (146 MB of traffic: 2.7 TB/s).
__global__ void __launch_bounds__(256) nchw_f32_to_nhwc16_bf16_x4_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                                         int n, int c, int h, int w, float scale) {
    const size_t plane4 = ((size_t)h * w) >> 2, total4 = (size_t)n * plane4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t img = i / plane4, q = i - img * plane4;
        float4 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = k < c ? __ldcs(reinterpret_cast<const float4*>(in + (img * c + k) * (plane4 << 2)) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint4* o = reinterpret_cast<uint4*>(out + (i << 2) * 16);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            uint32_t pk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float a = reinterpret_cast<const float*>(&v[2 * k])[px] * scale;
                const float b = reinterpret_cast<const float*>(&v[2 * k + 1])[px] * scale;
                const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
                pk[k] = *reinterpret_cast<const uint32_t*>(&hh);
            }
#ifndef PNNP_HOST_EMUL
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"       // one whole 32-byte sector per instruction
                         ::"l"(o + 2 * px), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
#else
            o[2 * px] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            o[2 * px + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
#endif
        }
    }
}

// 2x2 max pooling, NHWC bf16, 8 channels (16 bytes) per thread
__global__ void maxpool2x2_nhwc_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int n, int h,
                                            int w, int c) {
    const int ho = h / 2, wo = w / 2, c8 = c / 8;
    const size_t total = (size_t)n * ho * wo * c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c8);
        size_t r = i / c8;
        const int xo = (int)(r % wo); r /= wo;
        const int yo = (int)(r % ho);
        const int img = (int)(r / ho);
        const __nv_bfloat16* base = in + (((size_t)img * h + 2 * yo) * w + 2 * xo) * c + cc * 8;
        const uint4 q00 = *reinterpret_cast<const uint4*>(base);
        const uint4 q01 = *reinterpret_cast<const uint4*>(base + c);
        const uint4 q10 = *reinterpret_cast<const uint4*>(base + (size_t)w * c);
        const uint4 q11 = *reinterpret_cast<const uint4*>(base + (size_t)w * c + c);
        const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&q00);
        const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&q01);
        const __nv_bfloat162* cq = reinterpret_cast<const __nv_bfloat162*>(&q10);
        const __nv_bfloat162* d = reinterpret_cast<const __nv_bfloat162*>(&q11);
        uint4 o;
        __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) op[k] = __hmax2(__hmax2(a[k], b[k]), __hmax2(cq[k], d[k]));
        *reinterpret_cast<uint4*>(out + (((size_t)img * ho + yo) * wo + xo) * c + cc * 8) = o;
    }
}

}  // namespace pnnp
