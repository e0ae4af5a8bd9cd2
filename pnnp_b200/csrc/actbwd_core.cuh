// act_bwd_bias, second form (the default since r02; PNNP_ACTBWD_V2=0 for the first form): the same in-place  g *= act'(out)  and per-channel bias
// gradient as act_bwd_bias_kernel (train_kernels.cu), without the 64-bit `i % (c/8)` per 16-byte item that makes up more than half of
// that kernel's executed instructions: c/8 is a power of two and the grid stride is a multiple of it, so a thread's channel group
// is  start & (c/8 - 1)  once and for all; item indices are 32-bit.
// Written as PHASES separated by __syncthreads() so that the CPU suite can run phase after phase over the threads of a block
// (tests/emul/, test infrastructure only).
#pragma once
#include <cstddef>
#include <cstdint>
#ifndef PNNP_HOST_EMUL
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#endif

namespace pnnp {

constexpr int kAb2Threads = 256;

struct ActBwd2Args {
    uint16_t* g; const uint16_t* out;        // NHWC bf16 bit patterns
    uint32_t items;                          // pixels * (c / 8) 16-byte items
    int c, act_kind;                         // act_kind: 0 none (bias sums only), 1 LeakyReLU 0.2, 2 ReLU
};

__device__ __forceinline__ float ab2_bf(uint32_t h) { return __uint_as_float(h << 16); }
__device__ __forceinline__ uint32_t ab2_rn(float f) {                      // float -> bf16 bits, round to nearest even (finite inputs)
    const uint32_t u = __float_as_uint(f);
    return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}

// phase 1: clear the block's bias accumulators
__device__ __forceinline__ void ab2_clear(int tid, const ActBwd2Args& a, float* s_b) {
    for (int i = tid; i < a.c; i += kAb2Threads) s_b[i] = 0.f;
}

// phase 2: the thread's items (all of one channel group); returns nothing, adds its 8 sums to the block accumulators
template <typename AddFn>
__device__ __forceinline__ void ab2_main(int tid, uint32_t block, uint32_t nblocks, const ActBwd2Args& a, float* s_b, AddFn add) {
    const uint32_t c8m = (uint32_t)(a.c >> 3) - 1u;
    const uint32_t start = block * kAb2Threads + (uint32_t)tid, stride = nblocks * kAb2Threads;
    const uint32_t cg = start & c8m;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t i0 = start;
    if (a.act_kind == 0) {
        // bias sums only (read-only pass; 21 of a UNet step's launches): four independent 16-byte loads in flight per thread — with one
        // the pass was latency-bound at 2.8 TB/s (r02 launch list: 0.46 ms of the 4.4 ms step).  Same additions in the same order.
        for (; a.items > 3u * stride && i0 < a.items - 3u * stride; i0 += 4u * stride) {
            uint4 gv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) gv[u] = *reinterpret_cast<const uint4*>(a.g + (size_t)(i0 + (uint32_t)u * stride) * 8);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t* gw = reinterpret_cast<const uint32_t*>(&gv[u]);
#pragma unroll
                for (int k = 0; k < 4; ++k) { acc[2 * k] += ab2_bf(gw[k] & 0xFFFFu); acc[2 * k + 1] += ab2_bf(gw[k] >> 16); }
            }
        }
    }
    for (uint32_t i = i0; i < a.items; i += stride) {
        uint4 gv = *reinterpret_cast<const uint4*>(a.g + (size_t)i * 8);
        uint32_t* gw = reinterpret_cast<uint32_t*>(&gv);
        if (a.act_kind != 0) {
            const uint4 ov = *reinterpret_cast<const uint4*>(a.out + (size_t)i * 8);
            const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ov);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float x = ab2_bf(gw[k] & 0xFFFFu), y = ab2_bf(gw[k] >> 16);
                const float ox = ab2_bf(ow[k] & 0xFFFFu), oy = ab2_bf(ow[k] >> 16);
                if (a.act_kind == 1) { x *= ox > 0.f ? 1.f : 0.2f; y *= oy > 0.f ? 1.f : 0.2f; }
                else { x = ox > 0.f ? x : 0.f; y = oy > 0.f ? y : 0.f; }
                gw[k] = ab2_rn(x) | (ab2_rn(y) << 16);
            }
            *reinterpret_cast<uint4*>(a.g + (size_t)i * 8) = gv;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[2 * k] += ab2_bf(gw[k] & 0xFFFFu); acc[2 * k + 1] += ab2_bf(gw[k] >> 16); }
    }
    if (start < a.items)
#pragma unroll
        for (int k = 0; k < 8; ++k) add(&s_b[cg * 8 + k], acc[k]);
}

// phase 3: one global atomic per channel and block
template <typename AddFn>
__device__ __forceinline__ void ab2_flush(int tid, const ActBwd2Args& a, const float* s_b, float* dbias, AddFn add) {
    if (dbias) for (int i = tid; i < a.c; i += kAb2Threads) add(&dbias[i], s_b[i]);
}

}  // namespace pnnp
