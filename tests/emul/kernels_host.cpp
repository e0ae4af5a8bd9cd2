// The grid-stride kernels of pack.cu and crop_aug.cu compiled for the HOST and run thread by thread — TEST INFRASTRUCTURE ONLY
// (tests/test_device_kernels_on_cpu.py).  Same kernel source as the GPU build (csrc/pack_kernels.cuh, csrc/crop_kernels.cuh);
// the entry points mirror the C ABI's argument checks only as far as the tests need.
#define PNNP_HOST_EMUL 1
#include "cuda_host_shim.h"
#include "../../pnnp_b200/csrc/pack_kernels.cuh"
#include "../../pnnp_b200/csrc/crop_kernels.cuh"
#include "../../pnnp_b200/csrc/layout_kernels.cuh"
#include "../../pnnp_b200/csrc/ssim_core.cuh"
#include "../../pnnp_b200/csrc/copy_kernels.cuh"
#include "../../pnnp_b200/csrc/actbwd_core.cuh"

using namespace pnnp;

extern "C" {

// grid x block are free parameters: results must not depend on them
int emul_pack_norm_u16(const uint16_t* raw, float* out, int n, int H, int W, double wp, const double* black4, int norm, int clip,
                       int vec, int grid, int block) {
    const PackArgs a = make_pack_args(raw, out, n, H, W, wp, black4, norm, clip);
    if (vec) {
        if ((W / 2) % 4) return 1;
        if (a.use_rcp) EMUL_LAUNCH(grid, block, (pack_norm_kernel<uint16_t, true, true>(a)));      // as pack.cu selects
        else EMUL_LAUNCH(grid, block, (pack_norm_kernel<uint16_t, true>(a)));
    }
    else EMUL_LAUNCH(grid, block, (pack_norm_kernel<uint16_t, false>(a)));
    return 0;
}
int emul_pack_norm_f32(const float* raw, float* out, int n, int H, int W, double wp, const double* black4, int norm, int clip,
                       int vec, int grid, int block) {
    const PackArgs a = make_pack_args(raw, out, n, H, W, wp, black4, norm, clip);
    if (vec) {
        if ((W / 2) % 4) return 1;
        if (a.use_rcp) EMUL_LAUNCH(grid, block, (pack_norm_kernel<float, true, true>(a)));
        else EMUL_LAUNCH(grid, block, (pack_norm_kernel<float, true>(a)));
    }
    else EMUL_LAUNCH(grid, block, (pack_norm_kernel<float, false>(a)));
    return 0;
}
int emul_unpack_quant(const float* packed, uint16_t* raw, int n, int h, int w, float wp, float bl, int vec, int grid, int block) {
    const float span = wp - bl;
    if (vec) { if (w % 4) return 1; EMUL_LAUNCH(grid, block, (unpack_quant_kernel<true>(packed, raw, n, h, w, span, bl))); }
    else EMUL_LAUNCH(grid, block, (unpack_quant_kernel<false>(packed, raw, n, h, w, span, bl)));
    return 0;
}
int emul_pack_norm_dark_u16(const uint16_t* raw, const void* dark, int dark_is_f64, float* out, int n, int H, int W, double wp,
                            const double* b, int norm, int clip, double add_mean, int use_mean, double add_bias, int use_bias,
                            int grid, int block) {
    if (dark_is_f64)
        EMUL_LAUNCH(grid, block, (pack_norm_dark_kernel<double>(raw, static_cast<const double*>(dark), out, n, H, W, wp, b[0], b[1], b[2],
                                                                b[3], norm, clip, add_mean, use_mean, add_bias, use_bias)));
    else
        EMUL_LAUNCH(grid, block, (pack_norm_dark_kernel<float>(raw, static_cast<const float*>(dark), out, n, H, W, wp, b[0], b[1], b[2],
                                                               b[3], norm, clip, (float)add_mean, use_mean, (float)add_bias, use_bias)));
    return 0;
}

int emul_crop_aug(const float* frame, float* out, int c, int h, int w, int patch, int n, const int* hs, const int* ws, const int* mode,
                  int grid, int block) {
    if (n < 1 || n > kMaxCrops || (patch & 3)) return 1;
    CropArgs a{};
    a.frame = frame; a.out = out; a.c = c; a.h = h; a.w = w; a.patch = patch; a.n = n;
    for (int k = 0; k < n; ++k) { a.hs[k] = hs[k]; a.ws[k] = ws[k]; a.mode[k] = mode[k]; }
    EMUL_LAUNCH(grid, block, (crop_aug_kernel(a)));
    return 0;
}
static int geom(TileGeom& g, int c, int h, int w, int patch, int base) {       // crop_aug.cu: tile_geom
    if (c < 1 || base < 0 || (base & 1) || patch <= base || (patch & 3)) return 1;
    g.c = c; g.h = h; g.w = w; g.patch = patch; g.d = base / 2; g.l = patch - base;
    g.nh = h / g.l + 1; g.nw = w / g.l + 1;
    return (h + base < patch || w + base < patch || g.d >= h || g.d >= w) ? 1 : 0;
}
int emul_eval_crop(const float* frame, float* tiles, int c, int h, int w, int patch, int base, int grid, int block) {
    TileGeom g;
    if (geom(g, c, h, w, patch, base)) return 1;
    EMUL_LAUNCH(grid, block, (eval_crop_kernel(frame, tiles, g)));
    return 0;
}
int emul_eval_merge(const float* tiles, float* frame, int c, int h, int w, int patch, int base, int grid, int block) {
    TileGeom g;
    if (geom(g, c, h, w, patch, base)) return 1;
    EMUL_LAUNCH(grid, block, (eval_merge_kernel(tiles, frame, g)));
    return 0;
}
int emul_wb_gains(float* data, int n, int c, int h, int w, float rgb_gain, const int* kind, const double* gain, int grid, int block) {
    const size_t plane = (size_t)h * w;
    if (c > 8 || (plane & 3)) return 1;
    GainArgs a{};
    a.common = rgb_gain;
    for (int ch = 0; ch < c; ++ch) { a.kind[ch] = kind[ch]; a.g64[ch] = gain[ch]; a.g32[ch] = (float)gain[ch]; }
    EMUL_LAUNCH(grid, block, (wb_gains_kernel(data, plane / 4, c, (size_t)n * c * plane / 4, a)));
    return 0;
}

int emul_nchw_to_nhwc16(const float* in, uint16_t* out, int n, int c, int h, int w, float scale, int v2, int grid, int block) {
    if (c > 16) return 1;
    if (v2) { if (((size_t)h * w) & 3) return 1; EMUL_LAUNCH(grid, block, (nchw_f32_to_nhwc16_bf16_x4_kernel(in, reinterpret_cast<__nv_bfloat16*>(out), n, c, h, w, scale))); }
    else EMUL_LAUNCH(grid, block, (nchw_f32_to_nhwc16_bf16_kernel(in, reinterpret_cast<__nv_bfloat16*>(out), n, c, h, w, scale)));
    return 0;
}
int emul_maxpool2x2_nhwc(const uint16_t* in, uint16_t* out, int n, int h, int w, int c, int grid, int block) {
    if ((c % 8) || (h & 1) || (w & 1)) return 1;
    EMUL_LAUNCH(grid, block, (maxpool2x2_nhwc_bf16_kernel(reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), n, h, w, c)));
    return 0;
}

// ssim_mse_v2_kernel (eval_metrics.cu) phase by phase: every phase runs for all 256 threads of a block before the next one starts,
// which is what the __syncthreads() between them guarantees on the device.  sums: n x (3 + c) doubles as pnnp_eval_epilogue
// lays them out ([0], [1] = illuminance dots, supplied by the caller when use_gain).
int emul_ssim_mse_v2(const float* dn, const float* hr, int n, int c, int h, int w, float scale, int use_gain, double* sums) {
    static Ssim2Tile tile;
    const int stride = 3 + c;
    for (int plane = 0; plane < n * c; ++plane) {
        const int frame = plane / c, ch = plane - frame * c;
        Ssim2Args g{dn, hr, c, h, w, scale, 1.0f, use_gain};
        if (use_gain) g.gain = (float)sums[frame * stride + 0] / (float)sums[frame * stride + 1];
        for (int by = 0; by < (h + kS2TileY - 1) / kS2TileY; ++by)
            for (int bx = 0; bx < (w + kS2TileX - 1) / kS2TileX; ++bx) {
                const int x0 = bx * kS2TileX, y0 = by * kS2TileY;
                double se = 0.0, ssum = 0.0;
                for (int tid = 0; tid < kS2Threads; ++tid) se += ssim2_load(tid, g, plane, x0, y0, tile);
                for (int tid = 0; tid < kS2Threads; ++tid) ssim2_hsum(tid, tile);
                for (int tid = 0; tid < kS2Threads; ++tid) ssum += ssim2_vsum(tid, g, x0, y0, tile);
                sums[frame * stride + 2] += se;
                sums[frame * stride + 3 + ch] += ssum;
            }
    }
    return 0;
}

int emul_strided_copy_batch(const pnnp_copy_desc* descs, int n_desc, int v2, int grid, int block) {
    gridDim.y = (unsigned)n_desc;
    for (int y = 0; y < n_desc; ++y) {
        blockIdx.y = (unsigned)y;
        if (v2) EMUL_LAUNCH(grid, block, (strided_copy_batch_v2_kernel(descs)));
        else EMUL_LAUNCH(grid, block, (strided_copy_batch_kernel(descs)));
    }
    blockIdx.y = 0; gridDim.y = 1;
    return 0;
}

// act_bwd_bias_v2_kernel (train_kernels.cu) phase by phase; the atomics of the device become plain additions (threads run one after
// the other), so the bias sums agree with the device up to summation order and the in-place gradient update bit for bit
struct HostAdd { void operator()(float* p, float v) const { *p += v; } };
int emul_act_bwd_bias_v2(uint16_t* g, const uint16_t* out, float* dbias, uint32_t items, int c, int act_kind, int nblocks) {
    static float s_b[2048];
    if (c > 2048 || ((c / 8) & (c / 8 - 1))) return 1;
    ActBwd2Args a{g, out, items, c, act_kind};
    for (int b = 0; b < nblocks; ++b) {
        for (int t = 0; t < kAb2Threads; ++t) ab2_clear(t, a, s_b);
        for (int t = 0; t < kAb2Threads; ++t) ab2_main(t, (uint32_t)b, (uint32_t)nblocks, a, s_b, HostAdd());
        for (int t = 0; t < kAb2Threads; ++t) ab2_flush(t, a, s_b, dbias, HostAdd());
    }
    return 0;
}

}  // extern "C"
