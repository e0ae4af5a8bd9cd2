// Backward-pass helper kernels of the training step (launchers and the C ABI: train_kernels.cu).  They live in a header so that the
// CPU suite can compile this very source for the host and run whole CTAs of it — warp shuffles, shared-memory accumulators and
// block barriers included — on the lock-step fibre emulator (tests/emul/simt_host.h; test infrastructure only).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>
#ifndef PNNP_HOST_EMUL
#include <cuda_bf16.h>
#endif
#include "actbwd_core.cuh"

namespace pnnp {

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- L1 loss (F.l1_loss(pred.clamp(0,1), hr), mean reduction)
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ pred, const float* __restrict__ hr,
                                                      float* __restrict__ gpred, size_t total, float inv_total, double* loss_sum) {
    double acc = 0.0;
    // d|x|/dx = sign(x) (0 at 0, as torch); clamp passes the gradient only for 0 <= p <= 1 (torch.clamp semantics)
    auto one = [&](float p, float t) -> float {
        const float pc = fminf(fmaxf(p, 0.f), 1.f);
        const float d = pc - t;
        acc += (double)fabsf(d);
        const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        return (p >= 0.f && p <= 1.f) ? s * inv_total : 0.f;
    };
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(hr) | reinterpret_cast<uintptr_t>(gpred)) & 15) == 0;
    const size_t quads = vec ? total / 4 : 0;                                // 16-byte items; the (< 4 element) rest goes one by one
    for (size_t q = tid; q < quads; q += nthr) {
        const float4 p = reinterpret_cast<const float4*>(pred)[q], t = reinterpret_cast<const float4*>(hr)[q];
        float4 g;
        g.x = one(p.x, t.x); g.y = one(p.y, t.y); g.z = one(p.z, t.z); g.w = one(p.w, t.w);
        reinterpret_cast<float4*>(gpred)[q] = g;
    }
    for (size_t i = quads * 4 + tid; i < total; i += nthr) gpred[i] = one(pred[i], hr[i]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss_sum, acc);
}

// packed pairs of fp32 (FFMA2: the same IEEE fma on both halves, half the issue slots)
#ifndef PNNP_HOST_EMUL
__device__ __forceinline__ uint64_t hb_pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void hb_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t hb_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#else
static inline uint64_t hb_pack(float a, float b) { return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32); }
static inline void hb_unpack(uint64_t v, float& a, float& b) { a = __uint_as_float((uint32_t)v); b = __uint_as_float((uint32_t)(v >> 32)); }
static inline uint64_t hb_fma(uint64_t x, uint64_t y, uint64_t z) {
    float a, b, c, d, e, f; hb_unpack(x, a, b); hb_unpack(y, c, d); hb_unpack(z, e, f); return hb_pack(fmaf(a, c, e), fmaf(b, d, f));
}
#endif

// ---------------------------------------------------------------- 1x1 head backward (out_nc <= 4, cin <= 64)
// gpred: NCHW fp32 [n][co][h][w]; act: NHWC bf16 [n,h,w,cin] = LeakyReLU output feeding the head;
// gact (out): NHWC bf16 gradient w.r.t. the PRE-activation of that layer (already multiplied by act');
// dW[co][cin], db[co], dbias_prev[cin] (bias gradient of the layer feeding the head): fp32, accumulated with atomics.
// Work item = (pixel, group of 8 channels): one 16-byte load of the activation, one 16-byte store of the gradient; a thread
// keeps its channel group over all its pixels (grid * 256 is a multiple of cin / 8) and accumulates its 4 x 8 slice of dW, its 8
// bias sums (and db in group 0) in registers; lanes with the same group are combined by shuffles before the shared / global atomics.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ gpred, const __nv_bfloat16* __restrict__ act,
                                                       const float* __restrict__ W, __nv_bfloat16* __restrict__ gact,
                                                       float* dW, float* db, float* dbias_prev, int n, int h, int w, int cin,
                                                       int co, int act_kind) {
    extern __shared__ float s_acc[];                 // [co*cin] dW + [co] db + [cin] bias gradient of the previous conv
    float* s_dw = s_acc; float* s_db = s_acc + co * cin; float* s_dbp = s_db + co;
    for (int i = threadIdx.x; i < co * cin + co + cin; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int groups = cin / 8;                      // 1, 2, 4 or 8: divides 32, so a warp holds whole pixels
    int sh = 0;
    while ((1 << sh) < groups) ++sh;
    const size_t plane = (size_t)h * w, npix = (size_t)n * plane;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int grp = (int)(tid & (size_t)(groups - 1)), c = grp * 8;
    float adb[4] = {0.f, 0.f, 0.f, 0.f}, abp[8];
    uint64_t wv2[4][4], adw2[4][4];                   // channel pairs (2 k2, 2 k2 + 1): weights and the dW accumulators
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            wv2[o][k2] = o < co ? hb_pack(W[o * cin + c + 2 * k2], W[o * cin + c + 2 * k2 + 1]) : hb_pack(0.f, 0.f);
            adw2[o][k2] = hb_pack(0.f, 0.f);
        }
#pragma unroll
    for (int k = 0; k < 8; ++k) abp[k] = 0.f;
    const float slope = act_kind == 1 ? 0.2f : (act_kind == 2 ? 0.f : 1.f);
    // A thread's items are ps pixels apart (grid * 256 is a multiple of the group count), so (image, offset inside the plane) advance
    // by increment and carry: the two 64-bit divisions per item of the first form were more than half of its executed instructions
    // (r02 launch list: 157 us for 302 MB).  Two items per trip, loads first; the additions keep their order (same results bit for bit).
    // Pointers advance with the items (one add each; a carry adds the co - 1 planes between two images): the first form of this loop
    // rebuilt (img * co + o) * plane + off and pix * cin + c in 64 bits for every load and store — 87 of its 250 instructions per item.
    const size_t ps = ((size_t)gridDim.x * blockDim.x) >> sh, pe = ps * (size_t)cin, nelem = npix * (size_t)cin;
    size_t off, eo;                                   // offset inside the image plane; element offset of the item in act / gact
    const float* gptr;                                // gpred of (image, output 0, off)
    {
        const size_t pix = tid >> sh, img = pix / plane;
        off = pix - img * plane;
        eo = pix * (size_t)cin + (size_t)c;
        gptr = gpred + img * (size_t)co * plane + off;
    }
    const size_t carry = (size_t)(co - 1) * plane;
    auto advance = [&]() {
        eo += pe; off += ps; gptr += ps;
        while (off >= plane) { off -= plane; gptr += carry; }
    };
    auto fetch = [&](float (&gp)[4], uint4& av) {
#pragma unroll
        for (int o = 0; o < 4; ++o) gp[o] = o < co ? gptr[(size_t)o * plane] : 0.f;
        av = *reinterpret_cast<const uint4*>(act + eo);
    };
    auto item = [&](size_t at, const float (&gp)[4], const uint4& av) {
        const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
        uint32_t gw[4];
        uint64_t gp2[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) gp2[o] = hb_pack(gp[o], gp[o]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            const float a0 = __uint_as_float(aw[k2] << 16), a1 = __uint_as_float(aw[k2] & 0xFFFF0000u);
            const uint64_t a2 = hb_pack(a0, a1);
            uint64_t g2 = hb_pack(0.f, 0.f);
#pragma unroll
            for (int o = 0; o < 4; ++o) {             // the fmas of the scalar form, two channels per instruction (same order: same bits)
                g2 = hb_fma(gp2[o], wv2[o][k2], g2);
                adw2[o][k2] = hb_fma(gp2[o], a2, adw2[o][k2]);
            }
            float g0, g1;
            hb_unpack(g2, g0, g1);
            g0 *= a0 > 0.f ? 1.f : slope; g1 *= a1 > 0.f ? 1.f : slope;
            abp[2 * k2] += g0; abp[2 * k2 + 1] += g1;      // pre-activation gradient: what the previous conv's bias gradient sums
            const __nv_bfloat162 hb = __floats2bfloat162_rn(g0, g1);
            gw[k2] = *reinterpret_cast<const uint32_t*>(&hb);
        }
        if (grp == 0) {
#pragma unroll
            for (int o = 0; o < 4; ++o) adb[o] += gp[o];
        }
        *reinterpret_cast<uint4*>(gact + at) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
    };
    constexpr int kInFlight = 2;                     // items whose loads are issued before the first one is used (1 and two divisions per item: 157 us; 2: 113; 4 at 128 registers: 111)
    while (eo < nelem) {
        float gpv[kInFlight][4];
        uint4 avv[kInFlight];
        size_t at[kInFlight];
        bool on[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            at[u] = eo;
            on[u] = eo < nelem;
            if (on[u]) { fetch(gpv[u], avv[u]); advance(); }
        }
#pragma unroll
        for (int u = 0; u < kInFlight; ++u)
            if (on[u]) item(at[u], gpv[u], avv[u]);
    }
    float adw[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) hb_unpack(adw2[o][k2], adw[o][2 * k2], adw[o][2 * k2 + 1]);
    // lanes l, l + groups, l + 2 groups, ... hold the same channel group: combine them
    for (int o = groups; o < 32; o <<= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int k = 0; k < 8; ++k) adw[q][k] += __shfl_xor_sync(0xffffffffu, adw[q][k], o);
            adb[q] += __shfl_xor_sync(0xffffffffu, adb[q], o);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) abp[k] += __shfl_xor_sync(0xffffffffu, abp[k], o);
    }
    if ((int)(threadIdx.x & 31) < groups) {
#pragma unroll
        for (int o = 0; o < 4; ++o) if (o < co) {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(&s_dw[o * cin + c + k], adw[o][k]);
            if (grp == 0) atomicAdd(&s_db[o], adb[o]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s_dbp[c + k], abp[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < co * cin; i += blockDim.x) atomicAdd(&dW[i], s_dw[i]);
    for (int i = threadIdx.x; i < co; i += blockDim.x) atomicAdd(&db[i], s_db[i]);
    if (dbias_prev) for (int i = threadIdx.x; i < cin; i += blockDim.x) atomicAdd(&dbias_prev[i], s_dbp[i]);
}

// ---------------------------------------------------------------- activation backward + bias gradient
// g (in/out): NHWC bf16 gradient w.r.t. the activated output -> w.r.t. the pre-activation; out: the activated
// forward output (may be null for act == none); dbias[c] += sum over pixels of the pre-activation gradient.
__global__ void __launch_bounds__(256) act_bwd_bias_kernel(__nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ out,
                                                           float* dbias, size_t pixels, int c, int act_kind) {
    extern __shared__ float s_b[];                   // [c]
    for (int i = threadIdx.x; i < c; i += blockDim.x) s_b[i] = 0.f;
    __syncthreads();
    const int c8 = c / 8;                            // 16-byte groups per pixel
    const size_t total = pixels * c8;
    // a thread keeps the same channel group while striding over pixels when blockDim*gridDim % c8 == 0
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int my_cg = -1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cg = (int)(i % c8);
        if (cg != my_cg) {
            if (my_cg >= 0) for (int k = 0; k < 8; ++k) { atomicAdd(&s_b[my_cg * 8 + k], acc[k]); acc[k] = 0.f; }
            my_cg = cg;
        }
        uint4 gv = *reinterpret_cast<const uint4*>(g + i * 8);
        __nv_bfloat162* g2 = reinterpret_cast<__nv_bfloat162*>(&gv);
        if (act_kind != 0) {
            const uint4 ov = *reinterpret_cast<const uint4*>(out + i * 8);
            const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a = __low2float(g2[k]), b = __high2float(g2[k]);
                const float oa = __low2float(o2[k]), ob = __high2float(o2[k]);
                if (act_kind == 1) { a *= oa > 0.f ? 1.f : 0.2f; b *= ob > 0.f ? 1.f : 0.2f; }
                else { a = oa > 0.f ? a : 0.f; b = ob > 0.f ? b : 0.f; }
                g2[k] = __floats2bfloat162_rn(a, b);
            }
            *reinterpret_cast<uint4*>(g + i * 8) = gv;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[2 * k] += __low2float(g2[k]); acc[2 * k + 1] += __high2float(g2[k]); }
    }
    if (my_cg >= 0) for (int k = 0; k < 8; ++k) atomicAdd(&s_b[my_cg * 8 + k], acc[k]);
    __syncthreads();
    if (dbias) for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(&dbias[i], s_b[i]);
}

// ---------------------------------------------------------------- 2x2 max-pool backward (+ skip-connection gradient)
// gc[n,h,w,c] = (gskip ? gskip : 0) + gp[n,h/2,w/2,c] at the first arg-max of each window (row-major scan order, as torch),
// optionally times act'(cfull).  Work item = (pooled pixel, 8 channels): 16-byte loads / stores throughout.
// kBias: dbias[ch] += sum over pixels of the values written to gc (the bias gradient of the conv whose pre-activation gradient gc is) —
// the tensor is in registers here, the separate read-only pass (act_bwd_bias with act none) read it back from HBM / L2.  Needs
// grid * 256 to be a multiple of c / 8 (a thread keeps its channel group) and c floats of dynamic shared memory.
template <bool kBias>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ gp, const __nv_bfloat16* __restrict__ cfull,
                                                          const __nv_bfloat16* __restrict__ gskip, __nv_bfloat16* __restrict__ gc,
                                                          float* dbias, int n, int h, int w, int c, int act_kind) {
    extern __shared__ float s_pool_bias[];
    const int ho = h / 2, wo = w / 2, c8 = c / 8;
    const size_t total = (size_t)n * ho * wo * c8;
    const float sl = act_kind == 1 ? 0.2f : 0.f;
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (kBias) {
        for (int i = threadIdx.x; i < c; i += blockDim.x) s_pool_bias[i] = 0.f;
        __syncthreads();
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int cc, xo, yo, img;
        if (total <= 0xFFFFFFFFull) {                         // 32-bit index arithmetic (the 64-bit divisions were most of the kernel's instructions)
            uint32_t r = (uint32_t)i;
            const uint32_t q0 = r / (uint32_t)c8; cc = (int)(r - q0 * (uint32_t)c8); r = q0;
            const uint32_t q1 = r / (uint32_t)wo; xo = (int)(r - q1 * (uint32_t)wo); r = q1;
            const uint32_t q2 = r / (uint32_t)ho; yo = (int)(r - q2 * (uint32_t)ho); img = (int)q2;
        } else {
            size_t r = i / c8;
            cc = (int)(i % c8);
            xo = (int)(r % wo); r /= wo;
            yo = (int)(r % ho);
            img = (int)(r / ho);
        }
        const size_t base = (((size_t)img * h + 2 * yo) * w + 2 * xo) * c + cc * 8;
        const size_t idx[4] = {base, base + c, base + (size_t)w * c, base + (size_t)w * c + c};
        uint4 cv[4], sv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cv[k] = *reinterpret_cast<const uint4*>(cfull + idx[k]);
            sv[k] = gskip ? *reinterpret_cast<const uint4*>(gskip + idx[k]) : make_uint4(0u, 0u, 0u, 0u);
        }
        const uint4 gv = *reinterpret_cast<const uint4*>(gp + (((size_t)img * ho + yo) * wo + xo) * c + cc * 8);
        const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t ow[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                         // channel pair q of the group
            float v0[4], v1[4], s0[4], s1[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t cw = q == 0 ? cv[k].x : (q == 1 ? cv[k].y : (q == 2 ? cv[k].z : cv[k].w));
                const uint32_t sw = q == 0 ? sv[k].x : (q == 1 ? sv[k].y : (q == 2 ? sv[k].z : sv[k].w));
                v0[k] = __uint_as_float(cw << 16); v1[k] = __uint_as_float(cw & 0xFFFF0000u);
                s0[k] = __uint_as_float(sw << 16); s1[k] = __uint_as_float(sw & 0xFFFF0000u);
            }
            int a0 = 0, a1 = 0;
            float m0 = v0[0], m1 = v1[0];                     // the running maxima in registers (v0[a0] was a local-memory access)
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                if (v0[k] > m0) { m0 = v0[k]; a0 = k; }
                if (v1[k] > m1) { m1 = v1[k]; a1 = k; }
            }
            const float g0 = __uint_as_float(gw[q] << 16), g1 = __uint_as_float(gw[q] & 0xFFFF0000u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 o = __floats2bfloat162_rn(s0[k] + (k == a0 ? g0 : 0.f), s1[k] + (k == a1 ? g1 : 0.f));
                if (act_kind)           // fused act'(cfull): the sum is rounded to bf16 first, as when the two steps were separate kernels
                    o = __floats2bfloat162_rn(__low2float(o) * (v0[k] > 0.f ? 1.f : sl), __high2float(o) * (v1[k] > 0.f ? 1.f : sl));
                ow[k][q] = *reinterpret_cast<const uint32_t*>(&o);
                if (kBias) { bsum[2 * q] += __low2float(o); bsum[2 * q + 1] += __high2float(o); }   // the stored (bf16) values, as the separate pass summed
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(gc + idx[k]) = make_uint4(ow[k][0], ow[k][1], ow[k][2], ow[k][3]);
    }
    if (kBias) {
        const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int cc = (int)(tid % (size_t)c8);                // the channel group of every item of this thread
        if (tid < total) {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(&s_pool_bias[cc * 8 + k], bsum[k]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(&dbias[i], s_pool_bias[i]);
    }
}

// ---------------------------------------------------------------- Adam (torch.optim.Adam defaults: betas .9/.999, eps 1e-8, no decay)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, size_t total, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2, float gscale) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / sqrtf(bc2) + eps;          // torch: (sqrt(v) / sqrt(bias_correction2)) + eps
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

// Graph-friendly Adam: the step count and the learning rate live in device memory (state[0] = lr, state[1] = step as float), so a
// captured CUDA graph of the whole training step can be replayed while both change.  A one-thread kernel advances the step.
__global__ void adam_tick_kernel(float* state) { state[1] += 1.0f; }
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, size_t total, const float* __restrict__ state, float b1,
                                                       float b2, float eps, float gscale) {
    const float lr = state[0], t = state[1];
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

// act_bwd_bias, second form (actbwd_core.cuh): phases of one block, atomics on the shared / global accumulators
struct Ab2Atomic { __device__ __forceinline__ void operator()(float* p, float v) const { atomicAdd(p, v); } };
__global__ void __launch_bounds__(kAb2Threads) act_bwd_bias_v2_kernel(const ActBwd2Args a, float* dbias) {
    extern __shared__ float s_b2[];
    ab2_clear(threadIdx.x, a, s_b2);
    __syncthreads();
    ab2_main(threadIdx.x, blockIdx.x, gridDim.x, a, s_b2, Ab2Atomic());
    __syncthreads();
    ab2_flush(threadIdx.x, a, s_b2, dbias, Ab2Atomic());
}

}  // namespace pnnp
