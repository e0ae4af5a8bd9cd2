// D2 — batched random crop + 8-mode augmentation on the device
// (data_process/syn_datasets.py:69-107,162-173): crop = img[:, hs:hs+p, ws:ws+p];
// rot90 by (mode % 4) on the (H, W) axes (numpy.rot90, counter-clockwise), then a W-flip if mode // 4.
// A pure gather: one float4 of output per thread, source index computed per element.
#include "abi_common.h"
#include "../../include/pnnp_b200.h"

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
}  // namespace pnnp

using namespace pnnp;

extern "C" int pnnp_crop_aug(const float* frame, float* out, int c, int h, int w, int patch, int n,
                             const int* h_start_host, const int* w_start_host, const int* mode_host, void* stream) {
    if (!frame || !out || !h_start_host || !w_start_host || !mode_host) return fail("crop_aug: null pointer");
    if (n < 1 || n > kMaxCrops) return fail("crop_aug: 1..64 crops per call");
    if (patch < 4 || (patch & 3) || patch > h || patch > w) return fail("crop_aug: patch must be a multiple of 4 and fit the frame");
    CropArgs a{};
    a.frame = frame; a.out = out; a.c = c; a.h = h; a.w = w; a.patch = patch; a.n = n;
    for (int k = 0; k < n; ++k) {
        if (h_start_host[k] < 0 || w_start_host[k] < 0 || h_start_host[k] + patch > h || w_start_host[k] + patch > w)
            return fail("crop_aug: crop window outside the frame");
        if (mode_host[k] < 0 || mode_host[k] > 7) return fail("crop_aug: augmentation mode must be 0..7");
        a.hs[k] = h_start_host[k]; a.ws[k] = w_start_host[k]; a.mode[k] = mode_host[k];
    }
    const size_t total = (size_t)n * c * patch * (patch / 4);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    crop_aug_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
