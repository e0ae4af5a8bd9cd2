// Batched strided copy / cast kernels of train_kernels.cu (weight packing for the tensor-core layouts, gradient re-layout), kept in a
// header so that the CPU suite can compile these very kernels for the host and run them thread by thread (tests/emul/, test
// infrastructure only).
#pragma once
#include <cstddef>
#include <cstdint>
#ifndef PNNP_HOST_EMUL
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#endif
#include "../../include/pnnp_b200.h"

namespace pnnp {
__global__ void __launch_bounds__(256) strided_copy_batch_kernel(const pnnp_copy_desc* __restrict__ descs) {
    const pnnp_copy_desc d = descs[blockIdx.y];
    const long long n = (long long)d.dim[0] * d.dim[1] * d.dim[2] * d.dim[3];
    const float* src = static_cast<const float*>(d.src);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int i3 = (int)(r % d.dim[3]); r /= d.dim[3];
        const int i2 = (int)(r % d.dim[2]); r /= d.dim[2];
        const int i1 = (int)(r % d.dim[1]);
        const int i0 = (int)(r / d.dim[1]);
        const float v = src[i0 * d.sstride[0] + i1 * d.sstride[1] + i2 * d.sstride[2] + i3 * d.sstride[3]];
        const long long o = i0 * d.dstride[0] + i1 * d.dstride[1] + i2 * d.dstride[2] + i3 * d.dstride[3];
        if (d.dst_bf16) static_cast<__nv_bfloat16*>(d.dst)[o] = __float2bfloat16_rn(v);
        else static_cast<float*>(d.dst)[o] = v;
    }
}

// The default since r02 (4.88 -> 4.77 ms per training step; PNNP_COPY_V2=0 for the first form): the same copy with 32-bit index arithmetic (every descriptor has < 2^31 elements: checked
// by the launcher) — the three 64-bit divisions per element above make the kernel compute-bound (0.3 ms per training step for
// 31 MB in and out: 0.3-0.6 TB/s), and the innermost index is advanced without a division inside a thread's grid-stride walk when
// the stride is a multiple of dim[3].
__global__ void __launch_bounds__(256) strided_copy_batch_v2_kernel(const pnnp_copy_desc* __restrict__ descs) {
    const pnnp_copy_desc d = descs[blockIdx.y];
    const uint32_t d1 = (uint32_t)d.dim[1], d2 = (uint32_t)d.dim[2], d3 = (uint32_t)d.dim[3];
    const uint32_t n = (uint32_t)d.dim[0] * d1 * d2 * d3;
    const float* src = static_cast<const float*>(d.src);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t r = i;
        const uint32_t i3 = r % d3; r /= d3;
        const uint32_t i2 = r % d2; r /= d2;
        const uint32_t i1 = r % d1;
        const uint32_t i0 = r / d1;
        const float v = src[(long long)i0 * d.sstride[0] + (long long)i1 * d.sstride[1] + (long long)i2 * d.sstride[2] + (long long)i3 * d.sstride[3]];
        const long long o = (long long)i0 * d.dstride[0] + (long long)i1 * d.dstride[1] + (long long)i2 * d.dstride[2] + (long long)i3 * d.dstride[3];
        if (d.dst_bf16) static_cast<__nv_bfloat16*>(d.dst)[o] = __float2bfloat16_rn(v);
        else static_cast<float*>(d.dst)[o] = v;
    }
}
}  // namespace pnnp
