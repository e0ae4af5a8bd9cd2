// Backward-pass helper kernels of the synthetic-pair training step (T1: trainer_SID.py:93-101,
// losses/base_loss.py:92-103 — L1 loss on pred.clamp(0,1), Adam lr 1e-4).  The GEMM-shaped parts of
// the backward (dgrad, wgrad) run on the tcgen05 kernels (conv_tc.cu, wgrad_nhwc_tc.cu); everything here
// is element-wise / small-reduction work on CUDA cores:
//   l1_loss_kernel        loss = mean |clamp(pred,0,1) - hr| and d loss / d pred            (NCHW fp32)
//   head_bwd_kernel       backward of the 1x1 head conv10_1 (4 <- 32): data, weight and bias gradients
//   act_bwd_bias_kernel   g *= act'(out) in place (LeakyReLU 0.2 / ReLU / none) + per-channel bias gradient
//   maxpool_bwd_kernel    routes the pooled gradient to the arg-max of each 2x2 window (+ skip gradient)
//   adam_kernel           fused Adam over a flat fp32 parameter buffer
#include <cuda_bf16.h>
#include <cstdlib>
#include "abi_common.h"
#include "train_kernels.cuh"
#include "../../include/pnnp_b200.h"

namespace pnnp {

static int blocks_for(size_t items) { return (int)std::max<size_t>(1, std::min<size_t>((items + 255) / 256, 148 * 8)); }

}  // namespace pnnp

using namespace pnnp;

extern "C" int pnnp_l1_loss(const float* pred, const float* hr, float* gpred, size_t total, double* loss_sum, void* stream) {
    if (!pred || !hr || !gpred || !loss_sum || !total) return fail("l1_loss: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PNNP_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), st));
    l1_loss_kernel<<<blocks_for(total), 256, 0, st>>>(pred, hr, gpred, total, 1.0f / (float)total, loss_sum);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_head_bwd(const float* gpred, const void* act, const float* W, void* gact, float* dW, float* db,
                             float* dbias_prev, int n, int h, int w, int cin, int co, int act_kind, void* stream) {
    if (!gpred || !act || !W || !gact || !dW || !db || co < 1 || co > 4 || (cin != 8 && cin != 16 && cin != 32 && cin != 64))
        return fail("head_bwd: bad arguments (1..4 outputs, 8/16/32/64 input channels)");
    const size_t smem = sizeof(float) * (size_t)(co * cin + co + cin);
    // every block ends in 132 shuffles per thread and co * cin + co + cin same-address atomics: few, long-lived blocks — two per SM,
    // what 127 registers keep resident (r02, 8 crops: 296 blocks 100 us, 592 104, 1184 112, 2368 112; PNNP_HEADBWD_BLOCKS: timing experiments)
    static const int hb_max = getenv("PNNP_HEADBWD_BLOCKS") ? atoi(getenv("PNNP_HEADBWD_BLOCKS")) : 148 * 2;
    head_bwd_kernel<<<std::min(blocks_for((size_t)n * h * w * (cin / 8)), hb_max), 256, smem, (cudaStream_t)stream>>>(
        gpred, static_cast<const __nv_bfloat16*>(act), W, static_cast<__nv_bfloat16*>(gact), dW, db, dbias_prev, n, h, w, cin, co, act_kind);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_act_bwd_bias(void* g, const void* out, float* dbias, size_t pixels, int c, int act_kind, void* stream) {
    if (!g || (act_kind && !out) || (c % 8) || c > 1024) return fail("act_bwd_bias: bad arguments");
    // grid * 256 is a multiple of c/8 (c/8 is a power of two <= 128 for this network family): a thread keeps its channel group.
    // Every block ends with c global atomics, so small tensors get few blocks (>= 8 items per thread) instead of 1184 x c atomics.
    const size_t items = pixels * (size_t)(c / 8);
    static const int max_blocks = getenv("PNNP_ACTBWD_BLOCKS") ? atoi(getenv("PNNP_ACTBWD_BLOCKS")) : 148 * 4;   // r02: 592 blocks 4.32 ms per step, 1184 4.38, 148 4.47 (every block ends in c same-address atomics)
    int blocks = (int)std::max<size_t>(1, std::min<size_t>(items / (256 * 8), (size_t)max_blocks));
    const int quantum = std::max(1, (c / 8) / 256);                  // keep grid * 256 a multiple of c / 8
    blocks = std::max(quantum, blocks / quantum * quantum);
    const int c8 = c / 8;
    if (variant_on("PNNP_ACTBWD_V2") && (c8 & (c8 - 1)) == 0 && c8 <= 256 && items < (1ull << 32) - (1ull << 24)) {
        ActBwd2Args a{static_cast<uint16_t*>(g), static_cast<const uint16_t*>(out), (uint32_t)items, c, act_kind};
        act_bwd_bias_v2_kernel<<<blocks, kAb2Threads, sizeof(float) * c, (cudaStream_t)stream>>>(a, dbias);
        count_launch();
        PNNP_CUDA(cudaGetLastError());
        return 0;
    }
    act_bwd_bias_kernel<<<blocks, 256, sizeof(float) * c, (cudaStream_t)stream>>>(
        static_cast<__nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(out), dbias, pixels, c, act_kind);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_maxpool_bwd(const void* gp, const void* cfull, const void* gskip, void* gc, int n, int h, int w, int c, int act_kind,
                                void* stream) {
    if (!gp || !cfull || !gc || (h & 1) || (w & 1) || (c & 7)) return fail("maxpool_bwd: bad arguments (even h, w; c % 8 == 0)");
    maxpool_bwd_kernel<false><<<blocks_for((size_t)n * (h / 2) * (w / 2) * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16*>(gp), static_cast<const __nv_bfloat16*>(cfull), static_cast<const __nv_bfloat16*>(gskip),
        static_cast<__nv_bfloat16*>(gc), nullptr, n, h, w, c, act_kind);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_maxpool_bwd_bias(const void* gp, const void* cfull, const void* gskip, void* gc, float* dbias, int n, int h, int w,
                                     int c, int act_kind, void* stream) {
    if (!gp || !cfull || !gc || !dbias || (h & 1) || (w & 1) || (c & 7) || c > 2048 || (256 % (c / 8)))
        return fail("maxpool_bwd_bias: bad arguments (even h, w; c / 8 a divisor of 256)");
    // every block ends in c same-address atomics: at least four items per thread on small tensors (as pnnp_act_bwd_bias)
    const size_t items = (size_t)n * (h / 2) * (w / 2) * (c / 8);
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(items / (256 * 4), 148 * 8));
    maxpool_bwd_kernel<true><<<blocks, 256, sizeof(float) * c, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16*>(gp), static_cast<const __nv_bfloat16*>(cfull), static_cast<const __nv_bfloat16*>(gskip),
        static_cast<__nv_bfloat16*>(gc), dbias, n, h, w, c, act_kind);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_adam_step(float* p, const float* g, float* m, float* v, size_t total, float lr, float b1, float b2, float eps,
                              int step, float gscale, void* stream) {
    if (!p || !g || !m || !v || step < 1) return fail("adam_step: bad arguments");
    const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
    adam_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, total, lr, b1, b2, eps, bc1, bc2, gscale);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- batched strided copy / cast (weight packing, gradient layout)
// Every step the tensor-core kernels need the fp32 master weights in their bf16 [tap][rows][cin] layouts (forward, and the
// transposed + flipped data-gradient forms), and the weight gradients come out of the wgrad kernel as [tap][ci][co]: ~190 small
// permute / cast / copy launches when done with framework ops.  Here each of them is one descriptor (4 logical dims, source and
// destination strides in elements; a negative source stride expresses the 180-degree filter flip) and ONE launch walks them all.
#include "copy_kernels.cuh"      // strided_copy_batch_kernel, strided_copy_batch_v2_kernel

extern "C" int pnnp_adam_step_dev(float* p, const float* g, float* m, float* v, size_t total, float* state_dev, float b1, float b2,
                                  float eps, float gscale, void* stream) {
    if (!p || !g || !m || !v || !state_dev) return pnnp::fail("adam_step_dev: bad arguments");
    pnnp::adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state_dev);
    pnnp::adam_dev_kernel<<<pnnp::blocks_for(total), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, total, state_dev, b1, b2, eps, gscale);
    pnnp::count_launch(2);
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_strided_copy_batch(const pnnp_copy_desc* descs_dev, int n_desc, int blocks_per_desc, void* stream) {
    if (!descs_dev || n_desc < 1 || blocks_per_desc < 1) return pnnp::fail("strided_copy_batch: bad arguments");
    if (variant_on("PNNP_COPY_V2"))      // 32-bit index arithmetic; every descriptor the trainer builds is < 2^31 elements
        pnnp::strided_copy_batch_v2_kernel<<<dim3((unsigned)blocks_per_desc, (unsigned)n_desc), 256, 0, (cudaStream_t)stream>>>(descs_dev);
    else
    pnnp::strided_copy_batch_kernel<<<dim3((unsigned)blocks_per_desc, (unsigned)n_desc), 256, 0, (cudaStream_t)stream>>>(descs_dev);
    pnnp::count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
