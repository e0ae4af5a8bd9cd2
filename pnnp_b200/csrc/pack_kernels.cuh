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
    double span[4], rcp[4]; int use_rcp;         // wp - black per plane, its correctly rounded reciprocal (norm_one_d), both from the host
};

// host side: the per-plane divisor and its reciprocal; use_rcp only when the reciprocal form is proven equal to the division
inline PackArgs make_pack_args(const void* raw, float* out, int n, int H, int W, double wp, const double* black4, int norm, int clip) {
    PackArgs a{raw, out, n, H, W, wp, {black4[0], black4[1], black4[2], black4[3]}, norm, clip, {0, 0, 0, 0}, {0, 0, 0, 0}, 1};
    for (int c = 0; c < 4; ++c) {
        a.span[c] = wp - black4[c];
        a.rcp[c] = 1.0 / a.span[c];
        uint64_t bits;
        static_assert(sizeof(bits) == sizeof(double), "double is 64 bits");
        __builtin_memcpy(&bits, &a.span[c], 8);
        const uint64_t mant = bits & 0xFFFFFFFFFFFFFull, expo = (bits >> 52) & 0x7FF;
        // normal, finite divisor and reciprocal, significand not all ones, magnitudes far from the over- / underflow thresholds
        if (expo < 1023 - 400 || expo > 1023 + 400 || mant == 0xFFFFFFFFFFFFFull) a.use_rcp = 0;
    }
    return a;
}


#ifndef PNNP_PACK_ST
#define PNNP_PACK_ST __stcs                    // tools/ubench_pack.cu rebuilds the kernel with other store / load flavours
#endif
#ifndef PNNP_PACK_LD
#define PNNP_PACK_LD __ldcs
#endif
// Eight raw samples of one row: the load (kept apart from the conversion so that a thread can have several rows in flight) and
// the samples as doubles.
template <typename T> struct Load8;
template <> struct Load8<uint16_t> {
    struct Raw { uint4 q; };
    static __device__ __forceinline__ Raw ldraw(const uint16_t* p) { return Raw{PNNP_PACK_LD(reinterpret_cast<const uint4*>(p))}; }
    // the eight codes as doubles without a conversion instruction: 2^52 + n is the bit pattern (0x43300000, n); minus 2^52 is exact
    static __device__ __forceinline__ void cvt(const Raw& r, double (&v)[8]) {
        const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __dsub_rn(__hiloint2double(0x43300000, (int)(w[i] & 0xFFFFu)), 4503599627370496.0);
            v[2 * i + 1] = __dsub_rn(__hiloint2double(0x43300000, (int)(w[i] >> 16)), 4503599627370496.0);
        }
    }
};
template <> struct Load8<float> {
    struct Raw { float4 a, b; };
    static __device__ __forceinline__ Raw ldraw(const float* p) {
        return Raw{PNNP_PACK_LD(reinterpret_cast<const float4*>(p)), PNNP_PACK_LD(reinterpret_cast<const float4*>(p) + 1)};
    }
    static __device__ __forceinline__ void cvt(const Raw& r, double (&v)[8]) {
        v[0] = (double)r.a.x; v[1] = (double)r.a.y; v[2] = (double)r.a.z; v[3] = (double)r.a.w;
        v[4] = (double)r.b.x; v[5] = (double)r.b.y; v[6] = (double)r.b.z; v[7] = (double)r.b.w;
    }
};

// One vector item = 8 samples of an even raw row and the 8 below them -> one float4 in each of the four planes.
template <typename T, bool RCP, typename I>
struct PackItem {
    I f, y, x4;
    typename Load8<T>::Raw e, o;
    __device__ __forceinline__ void load(const PackArgs& a, I i, I w4, I h) {
        x4 = i % w4;
        const I t = i / w4;
        y = t % h;
        f = t / h;
        const T* r0 = static_cast<const T*>(a.raw) + ((size_t)f * a.H + 2 * (size_t)y) * a.W + 8 * (size_t)x4;
        e = Load8<T>::ldraw(r0);
        o = Load8<T>::ldraw(r0 + a.W);
    }
    __device__ __forceinline__ void finish(const PackArgs& a, size_t plane, int w) const {
        double ev[8], ov[8];
        Load8<T>::cvt(e, ev);
        Load8<T>::cvt(o, ov);
        // plane order R(0,0) G1(0,1) B(1,1) G2(1,0): even row -> R (even columns), G1 (odd); odd row -> G2 (even), B (odd)
#define PNNP_NORM(v, c) norm_one_d<RCP>(v, a.black[c], a.wp, a.span[c], a.rcp[c], a.norm, a.clip)
        float4 R, G1, B, G2;
        R.x = PNNP_NORM(ev[0], 0); R.y = PNNP_NORM(ev[2], 0); R.z = PNNP_NORM(ev[4], 0); R.w = PNNP_NORM(ev[6], 0);
        G1.x = PNNP_NORM(ev[1], 1); G1.y = PNNP_NORM(ev[3], 1); G1.z = PNNP_NORM(ev[5], 1); G1.w = PNNP_NORM(ev[7], 1);
        B.x = PNNP_NORM(ov[1], 2); B.y = PNNP_NORM(ov[3], 2); B.z = PNNP_NORM(ov[5], 2); B.w = PNNP_NORM(ov[7], 2);
        G2.x = PNNP_NORM(ov[0], 3); G2.y = PNNP_NORM(ov[2], 3); G2.z = PNNP_NORM(ov[4], 3); G2.w = PNNP_NORM(ov[6], 3);
#undef PNNP_NORM
        float* ob = a.out + (size_t)f * 4 * plane + (size_t)y * w + 4 * (size_t)x4;
        PNNP_PACK_ST(reinterpret_cast<float4*>(ob), R);
        PNNP_PACK_ST(reinterpret_cast<float4*>(ob + plane), G1);
        PNNP_PACK_ST(reinterpret_cast<float4*>(ob + 2 * plane), B);
        PNNP_PACK_ST(reinterpret_cast<float4*>(ob + 3 * plane), G2);
    }
};

// Two items per trip: both items' four 128-bit loads are issued before the first conversion (one item per trip kept 32 bytes per
// thread in flight: 3.6 TB/s); I = index type (32-bit whenever the item count allows: the two divisions per item were 64-bit).
template <typename T, bool RCP, typename I>
__device__ __forceinline__ void pack_vec_loop(const PackArgs& a, I total, I w4, I h, size_t plane, int w) {
    const I stride = (I)gridDim.x * (I)blockDim.x;
    I i = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x;
    for (; i < total && total - i > stride; i += 2 * stride) {
        PackItem<T, RCP, I> p0, p1;
        p0.load(a, i, w4, h);
        p1.load(a, i + stride, w4, h);
        p0.finish(a, plane, w);
        p1.finish(a, plane, w);
    }
    if (i < total) {
        PackItem<T, RCP, I> p0;
        p0.load(a, i, w4, h);
        p0.finish(a, plane, w);
    }
}

// RCP: the reciprocal form of the normalisation (norm_one_d; the launcher passes a.use_rcp) — vector path only
template <typename T, bool VEC, bool RCP = false>
__global__ void __launch_bounds__(256) pack_norm_kernel(const PackArgs a) {
    const T* raw = static_cast<const T*>(a.raw);
    const int h = a.H / 2, w = a.W / 2;
    const size_t plane = (size_t)h * w;
    if (VEC) {
        const int w4 = w / 4;                                  // float4 groups per packed row
        const size_t total = (size_t)a.n * h * w4;
        if (total < 0x40000000u) pack_vec_loop<T, RCP, uint32_t>(a, (uint32_t)total, (uint32_t)w4, (uint32_t)h, plane, w);
        else pack_vec_loop<T, RCP, size_t>(a, total, (size_t)w4, (size_t)h, plane, w);
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
