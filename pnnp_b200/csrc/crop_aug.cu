// D2 — batched random crop + 8-mode augmentation on the device
// (data_process/syn_datasets.py:69-107,162-173): crop = img[:, hs:hs+p, ws:ws+p];
// rot90 by (mode % 4) on the (H, W) axes (numpy.rot90, counter-clockwise), then a W-flip if mode // 4.
// A pure gather: one float4 of output per thread, source index computed per element.
#include <algorithm>
#include "abi_common.h"
#include "../../include/pnnp_b200.h"

#include "crop_kernels.cuh"

namespace pnnp {
static int tile_geom(TileGeom& g, int c, int h, int w, int patch, int base) {
    if (c < 1 || base < 0 || (base & 1) || patch <= base || (patch & 3)) return fail("eval tiling: patch must be a multiple of 4 and > base (even)");
    g.c = c; g.h = h; g.w = w; g.patch = patch; g.d = base / 2; g.l = patch - base;
    g.nh = h / g.l + 1; g.nw = w / g.l + 1;
    if (h + base < patch || w + base < patch) return fail("eval tiling: the padded frame is smaller than one tile");
    if (g.d >= h || g.d >= w) return fail("eval tiling: reflect padding needs base/2 < h, w");
    return 0;
}

}  // namespace pnnp

using namespace pnnp;

extern "C" int pnnp_wb_gains(float* data, int n, int c, int h, int w, float rgb_gain, const int* kind_host,
                             const double* gain_host, void* stream) {
    if (!data || !kind_host || !gain_host) return fail("wb_gains: null pointer");
    if (n < 1 || c < 1 || c > 8 || h < 1 || w < 1) return fail("wb_gains: 1..8 planes per crop");
    const size_t plane = (size_t)h * w;
    if (plane & 3) return fail("wb_gains: h * w must be a multiple of 4");
    GainArgs a{};
    a.common = rgb_gain;
    for (int ch = 0; ch < c; ++ch) {
        if (kind_host[ch] < 0 || kind_host[ch] > 2) return fail("wb_gains: kind must be 0 (none), 1 (float32) or 2 (float64)");
        a.kind[ch] = kind_host[ch]; a.g64[ch] = gain_host[ch]; a.g32[ch] = (float)gain_host[ch];
    }
    const size_t total4 = (size_t)n * c * plane / 4;
    wb_gains_kernel<<<(int)std::min<size_t>((total4 + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(data, plane / 4, c, total4, a);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_eval_crop(const float* frame, float* tiles, int c, int h, int w, int patch, int base, void* stream) {
    if (!frame || !tiles) return fail("eval_crop: null pointer");
    TileGeom g;
    if (int e = tile_geom(g, c, h, w, patch, base)) return e;
    const size_t total = (size_t)g.nh * g.nw * c * patch * (patch / 4);
    eval_crop_kernel<<<(int)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(frame, tiles, g);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_eval_merge(const float* tiles, float* frame, int c, int h, int w, int patch, int base, void* stream) {
    if (!frame || !tiles) return fail("eval_merge: null pointer");
    TileGeom g;
    if (int e = tile_geom(g, c, h, w, patch, base)) return e;
    if (h < g.l || w < g.l) return fail("eval_merge: frame smaller than one tile interior");
    const size_t total = (size_t)c * h * w;
    eval_merge_kernel<<<(int)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(tiles, frame, g);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

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
