// Bayer pack + black/white-level normalisation (P1) and its inverse (P2) for sm_100a.
//
// P1 raw2bayer  utils/isp_ops.py:84-96   uint16|float32 H x W  ->  float32 4 x H/2 x W/2
//    plane order R(0,0) G1(0,1) B(1,1) G2(1,0); (x - black_c) / (wp - black_c) in float64,
//    rounded to float32 once (the reference's `bias + bl` is an int64 array).
// P2 bayer2raw  utils/isp_ops.py:98-112  clip[0,1]; x*(wp-bl)+bl in float32 (two roundings);
//    truncating cast to uint16.
//
// Both are pure streaming kernels (HBM bound: 6 B per raw sample).  Vector path: each thread
// reads 8 samples of an even raw row and 8 of the odd row below (2 x 128-bit loads for uint16,
// 4 for float32) and writes one float4 into each of the four planes.
#include "abi_common.h"
#include "pack_kernels.cuh"
#include "../../include/pnnp_b200.h"

namespace pnnp {

static int grid_for(size_t work_items) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t want = (work_items + 255) / 256;
    return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sms * 8));
}

template <typename T>
static int pack_impl(const T* raw, float* out, int n, int H, int W, double wp, const double* black4, int norm,
                     int clip, void* stream) {
    if (!raw || !out || !black4) return fail("pack_norm: null pointer");
    if (n <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return fail("pack_norm: H and W must be positive and even");
    const PackArgs a = make_pack_args(raw, out, n, H, W, wp, black4, norm, clip);
    const int w = W / 2;
    const size_t row_bytes = (size_t)W * sizeof(T);
    const bool vec = (w % 4 == 0) && (reinterpret_cast<uintptr_t>(raw) % 16 == 0) && (row_bytes % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec && a.use_rcp) pack_norm_kernel<T, true, true><<<grid_for((size_t)n * (H / 2) * (w / 4)), 256, 0, st>>>(a);
    else if (vec) pack_norm_kernel<T, true><<<grid_for((size_t)n * (H / 2) * (w / 4)), 256, 0, st>>>(a);
    else pack_norm_kernel<T, false><<<grid_for((size_t)n * H * W), 256, 0, st>>>(a);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
}  // namespace pnnp

using namespace pnnp;

extern "C" int pnnp_pack_norm_u16(const uint16_t* raw, float* out, int n, int H, int W, double wp,
                                  const double* black4_host, int norm, int clip, void* stream) {
    return pack_impl<uint16_t>(raw, out, n, H, W, wp, black4_host, norm, clip, stream);
}
extern "C" int pnnp_pack_norm_f32(const float* raw, float* out, int n, int H, int W, double wp,
                                  const double* black4_host, int norm, int clip, void* stream) {
    return pack_impl<float>(raw, out, n, H, W, wp, black4_host, norm, clip, stream);
}
extern "C" int pnnp_pack_norm_dark_u16(const uint16_t* raw, const void* dark, int dark_is_f64, float* out, int n, int H, int W, double wp,
                                       const double* black4_host, int norm, int clip, double add_mean, int use_mean,
                                       double add_bias, int use_bias, void* stream) {
    if (!raw || !dark || !out || !black4_host) return fail("pack_norm_dark: null pointer");
    if (n <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return fail("pack_norm_dark: H and W must be positive and even");
    const int blocks = grid_for((size_t)n * H * W);
    cudaStream_t st = (cudaStream_t)stream;
    const double* b = black4_host;
    if (dark_is_f64)
        pack_norm_dark_kernel<double><<<blocks, 256, 0, st>>>(raw, static_cast<const double*>(dark), out, n, H, W, wp, b[0], b[1], b[2], b[3],
                                                              norm, clip, add_mean, use_mean, add_bias, use_bias);
    else
        pack_norm_dark_kernel<float><<<blocks, 256, 0, st>>>(raw, static_cast<const float*>(dark), out, n, H, W, wp, b[0], b[1], b[2], b[3],
                                                             norm, clip, (float)add_mean, use_mean, (float)add_bias, use_bias);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int pnnp_unpack_quant(const float* packed, uint16_t* raw, int n, int h, int w, float wp, float bl,
                                 void* stream) {
    if (!packed || !raw) return fail("unpack_quant: null pointer");
    if (n <= 0 || h <= 0 || w <= 0) return fail("unpack_quant: empty shape");
    const float span = wp - bl;       // python ints in the reference: exact in float32 for sensor ranges
    const bool vec = (w % 4 == 0) && (reinterpret_cast<uintptr_t>(packed) % 16 == 0) && (reinterpret_cast<uintptr_t>(raw) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) unpack_quant_kernel<true><<<grid_for((size_t)n * h * (w / 4)), 256, 0, st>>>(packed, raw, n, h, w, span, bl);
    else unpack_quant_kernel<false><<<grid_for((size_t)n * 4 * h * w), 256, 0, st>>>(packed, raw, n, h, w, span, bl);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
