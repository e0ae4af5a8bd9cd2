/*
 * pnnp_b200 — C ABI of the B200-native PNNP hot path (libpnnp_b200.so).
 *
 * The reference (fenghansen/PNNP) is pure Python and has no FFI; each entry point below
 * names the reference function (file:line under the reference tree) it replaces.  The
 * reference-side binding is a ctypes stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless the name ends in
 *     `_host`; the library never allocates or frees caller buffers;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream);
 *   - return value 0 = ok, non-zero = error, message via pnnp_last_error() (thread-local);
 *   - RNG state is (seed, offset), passed in; nothing is hidden in the library;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns an error.
 */
#ifndef PNNP_B200_H
#define PNNP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNNP_ABI_VERSION 1

/* noise_code letters of generate_noisy_obs (data_process/process.py:598-603) as bits */
#define PNNP_CODE_P 0x01u /* 'p' Poisson shot noise (else Gaussian approximation)   */
#define PNNP_CODE_G 0x02u /* 'g' Tukey-lambda read noise (else Gaussian sigGs)       */
#define PNNP_CODE_R 0x04u /* 'r' per-(channel,row) banding noise                      */
#define PNNP_CODE_Q 0x08u /* 'q' quantisation noise                                   */
#define PNNP_CODE_D 0x10u /* 'd' per-channel bias                                     */
#define PNNP_CODE_B 0x20u /* 'b' black frame: read/row/q/bias all zero                */
/* hint (not a reference letter): every table row has flags == PNNP_F_K64|PNNP_F_SIG64, i.e. all
 * crops use sample_params-style np.float64 parameters, and |lam| >= 1e-3 (every camera of the
 * reference: process.py:215-309); lets the launcher pick the branch-free kernel */
#define PNNP_CODE_UNIFORM_F64 0x100u

/* arithmetic chain (which reference function's rounding sequence is reproduced) */
#define PNNP_CHAIN_NUMPY 0 /* generate_noisy_obs   process.py:591-631 (NEP-50 promotion)   */
#define PNNP_CHAIN_TORCH 1 /* generate_noisy_torch process.py:634-673 (float32 throughout) */

/* per-crop flags: which python scalars were np.float64 ("strong") in the reference call */
#define PNNP_F_K64 0x1u     /* K is np.float64 (sample_params)  → shot term is float64   */
#define PNNP_F_RATIO64 0x2u /* ratio is np.float64                                         */
#define PNNP_F_SIG64 0x4u   /* sigR is np.float64                                          */

/* One row per crop (128 bytes).  Mirrors the dict returned by sample_params /
 * sample_params_max (process.py:311-412): keys K sigTL sigR sigGs bias lam q ratio wp bl. */
typedef struct pnnp_noise_params {
    double K;       /* system gain                                  */
    double sigTL;   /* Tukey-lambda scale                           */
    double sigGs;   /* Gaussian read-noise sigma                    */
    double sigR;    /* row-noise sigma                              */
    double lam;     /* Tukey-lambda shape                           */
    double q;       /* quantisation step (used by the torch chain)  */
    double ratio;   /* exposure ratio                               */
    double span;    /* wp - bl                                      */
    double clip_lo; /* -bl / wp  (process.py:627)                   */
    double bias[4]; /* per-channel bias ('d')                       */
    uint32_t flags; /* PNNP_F_*                                     */
    uint32_t reserved[5];
} pnnp_noise_params;

const char* pnnp_last_error(void);
int pnnp_abi_version(void);
/* number of kernels this library has launched in the calling process (for gpu_launches) */
uint64_t pnnp_launch_count(void);
/* adds n to that counter: the caller replayed a CUDA graph holding n of this library's kernels */
void pnnp_count_graph_launches(uint64_t n);

/* P1 — raw2bayer(raw, wp, bl, norm, clip, bias)                 utils/isp_ops.py:84-96
 * raw: n frames of H x W (uint16 or float32), out: n x 4 x H/2 x W/2 float32, plane order
 * R(0,0) G1(0,1) B(1,1) G2(1,0).  black4_host[c] = bl + bias[c]; arithmetic in float64,
 * rounded to float32 once, exactly like the reference. */
int pnnp_pack_norm_u16(const uint16_t* raw, float* out, int n, int H, int W, double wp,
                       const double* black4_host, int norm, int clip, void* stream);
int pnnp_pack_norm_f32(const float* raw, float* out, int n, int H, int W, double wp,
                       const double* black4_host, int norm, int clip, void* stream);
/* Dark-shading correction fused into P1 (data_process/real_datasets.py:360-372 -> raw2bayer): per sample
 * v = raw - dark[y][x] [+ add_mean] [+ add_bias] in the dark map's precision (H x W float32, or float64 when
 * dark_is_f64 — NumPy's promotion of `uint16 array - map`), cast to float32, then the normalisation above. */
int pnnp_pack_norm_dark_u16(const uint16_t* raw, const void* dark, int dark_is_f64, float* out, int n, int H, int W, double wp,
                            const double* black4_host, int norm, int clip, double add_mean, int use_mean,
                            double add_bias, int use_bias, void* stream);

/* P2 — bayer2raw(packed, wp, bl)                                utils/isp_ops.py:98-112
 * packed: n x 4 x h x w float32 → raw: n x 2h x 2w uint16 (truncating cast). */
int pnnp_unpack_quant(const float* packed, uint16_t* raw, int n, int h, int w, float wp, float bl,
                      void* stream);

/* N1-N3 / N4 — fused noise synthesis, Philox4x32-10 draws generated in-kernel.
 * clean, noisy: n x c x h x w float32 (may alias);  table: n rows (device).
 * clip: the reference's `clip` argument (0 → clip to [-bl/wp, 1], non-zero → [0, 1]).
 * post_lo/post_hi: the caller's follow-up clamp, fused (syn_datasets.py:339-342,
 * trainer_SID.py:481-485); pass -INF/+INF for none.
 * crop i uses Philox key = seed, counter = (element index, stream, offset + i-independent) so
 * results do not depend on grid shape or on how crops are sharded across GPUs:
 * `crop_id0` is the global index of the first crop in this call. */
int pnnp_noise_synth(const float* clean, float* noisy, const pnnp_noise_params* table, int n, int c,
                     int h, int w, uint32_t code_bits, int chain, int ori, int clip, float post_lo,
                     float post_hi, uint64_t seed, uint64_t offset, uint64_t crop_id0, void* stream);

/* Same launch, but also writes the draws it used (any pointer may be NULL):
 * d_shot  n*c*h*w f32 — Poisson counts ('p') or the standard-normal shot draw,
 * d_read  n*c*h*w f32 — read-noise sample in DN (already scaled),
 * d_rowz  n*c*h   f32 — standard-normal row draw,
 * d_q     n*c*h*w f64 — numpy chain: U(-0.5,0.5) in DN; torch chain: U[0,1). */
int pnnp_noise_synth_debug(const float* clean, float* noisy, const pnnp_noise_params* table, int n,
                           int c, int h, int w, uint32_t code_bits, int chain, int ori, int clip,
                           float post_lo, float post_hi, uint64_t seed, uint64_t offset,
                           uint64_t crop_id0, float* d_shot, float* d_read, float* d_rowz,
                           double* d_q, void* stream);

/* Replay mode: identical arithmetic core, draws supplied by the caller (the reference's own
 * draws in the parity tests) → bit-exact against generate_noisy_obs / generate_noisy_torch. */
int pnnp_noise_synth_replay(const float* clean, float* noisy, const pnnp_noise_params* table, int n,
                            int c, int h, int w, uint32_t code_bits, int chain, int ori, int clip,
                            float post_lo, float post_hi, const float* d_shot, const float* d_read,
                            const float* d_rowz, const double* d_q, void* stream);

/* ------------------------------------------------------------------------------------------
 * U1 / U2 building blocks — UNetSeeInDark.forward (archs/Unet.py:54-99) and ResUnet.forward
 * (archs/ResUnet.py:46-88, archs/modules.py:130-197) as tcgen05/TMEM implicit-GEMM layers.
 * Activations: NHWC bf16, channel counts multiples of 16.  Weights: bf16 [taps][rows][cin_total]
 * (cin innermost), taps = 9 (ky*3+kx) | 1 | 4 (a*2+b for ConvTranspose2d(2, stride 2)).
 * ------------------------------------------------------------------------------------------ */
#define PNNP_CONV3 0 /* nn.Conv2d(k=3, s=1, p=1)                    */
#define PNNP_CONV1 1 /* nn.Conv2d(k=1)                              */
#define PNNP_CONVT 2 /* nn.ConvTranspose2d(k=2, s=2): output 2h x 2w */
#define PNNP_CONV3S2 3 /* nn.Conv2d(k=3, s=2, p=1): output h/2 x w/2 (ResUnet down-sampling, modules.py:130-138) */
/* 3x3 s1 p1 with the x-shift folded into N: weights [ky][kx*cout + co][cin], MMA N = 3*cout, the three
 * kx partial sums are combined across neighbouring pixels in the epilogue (cout <= 80) */
#define PNNP_CONV3X 4
#define PNNP_CONV3B 6 /* nn.Conv2d(k=3, p=1), single narrow source: all nine taps from ONE haloed TMA box (8 x 16 tiles, taps as descriptor offsets); weights [ky*3+kx][cout][cin] like PNNP_CONV3 */
#define PNNP_CONV2S2 5 /* nn.Conv2d(k=2, s=2, p=0): the data gradient of ConvTranspose2d(2, stride 2); weights [a*2+b][cout][cin] */
#define PNNP_ACT_NONE 0
#define PNNP_ACT_LEAKY02 1 /* nn.LeakyReLU(0.2)  Unet.py:52    */
#define PNNP_ACT_RELU 2    /* nn.ReLU            ResUnet.py:44 */
#define PNNP_OUT_NHWC_BF16 0
#define PNNP_OUT_NCHW_F32 1 /* network output: fp32 planes straight from the fp32 accumulators */

/* One layer.  in1/cin1: optional second K source = the skip tensor of torch.cat([up, skip], 1)
 * (Unet.py:72,77,82,87) — the concat is never materialised.  resid: optional NHWC bf16 tensor added
 * after the activation (ResidualBlock `output += short_cut(x)`, modules.py:193-197); resid_nchw:
 * optional fp32 NCHW tensor added to an OUT_NCHW_F32 output (`out = conv10 + x`, Unet.py:96).
 * h, w are the INPUT spatial size.  w_rows = rows allocated per tap in `weight` (>= cout, padded
 * with zero rows up to a multiple of 16). */
int pnnp_conv2d_tc(int mode, const void* in0, int cin0, const void* in1, int cin1, const void* weight,
                   int w_rows, const float* bias, void* out, int cout, int cout_stride, int n, int h,
                   int w, int act, int out_mode, const void* resid, const float* resid_nchw,
                   void* stream);
/* Descriptor form with the fused epilogue options.  pool_out: also write nn.MaxPool2d(2) of the output
 * (NHWC bf16, h/2 x w/2; Unet.py:57-69).  head_*: fuse a following 1x1 conv with <= 4 output channels
 * (conv10_1, Unet.py:93-98) — fp32 weights [head_cout][cout], output NCHW fp32 (+ resid_nchw); `out` may
 * then be NULL so the intermediate activation is never written. */
typedef struct pnnp_conv_desc {
    int mode, act, out_mode;
    int n, h, w;
    const void* in0; int cin0;
    const void* in1; int cin1;
    const void* weight; int w_rows;
    const float* bias;
    void* out; int cout; int cout_stride;
    const void* resid;
    const float* resid_nchw;
    void* pool_out;
    const float* head_w; const float* head_b; float* head_out; int head_cout;
    const void* mask; float mask_slope;   /* NHWC bf16 activation shaped like the output: out *= (mask > 0 ? 1 : mask_slope) — the
                                           * activation derivative fused into a data-gradient conv (training); NULL = off */
    int io_f32;                           /* 1: fp32-storage variant — in0 / in1 / weight / out / resid / pool_out are fp32 (same NHWC /
                                           * [tap][rows][cin] layouts), the MMAs are tcgen05 kind::tf32.  The north_star's "fp32" accuracy
                                           * class (bf16 reported separately); inference layers of UNetSeeInDark (3x3, x-mode, transposed,
                                           * 1x1, fused pool / head, two sources).  0: bf16 storage (default). */
} pnnp_conv_desc;
int pnnp_conv2d_tc_ex(const pnnp_conv_desc* desc, void* stream);
/* Non-zero if a tcgen05/TMA pipeline wait timed out since the last call (the kernels terminate
 * instead of hanging); synchronises the device. */
int pnnp_conv_pipeline_error(void);
/* network input: NCHW fp32 (c <= 16) * scale -> NHWC bf16 zero-padded to 16 channels */
int pnnp_nchw_to_nhwc16(const float* in, void* out, int n, int c, int h, int w, float scale,
                        void* stream);
/* the same conversion with fp32 output (input of the fp32-storage / tf32 variant) */
int pnnp_nchw_to_nhwc16_f32(const float* in, float* out, int n, int c, int h, int w, void* stream);
/* First layer fused with the pack boundary (archs/Unet.py:55 conv1_1, archs/ResUnet.py conv_in, fed by the packed planes that
 * data_process/process.py:625-631 / utils/isp_ops.py:84-96 produce): NCHW fp32 input with cin <= 4 channels -> conv3x3 pad 1
 * (weight: the module's own fp32 [cout][cin][3][3], rounded to bf16 in the kernel) + bias + activation (act: 0 none, 1 LeakyReLU
 * 0.2, 2 ReLU) -> NHWC bf16 with cout in {16, 32, 48, 64} channels.  One launch instead of pnnp_nchw_to_nhwc16 + a 16-channel conv. */
int pnnp_conv_first_nchw(const float* in, const float* weight, const float* bias, void* out, int n, int cin, int h, int w,
                         int cout, int act, void* stream);
int pnnp_conv_first_pipeline_error(void);
/* nn.MaxPool2d(2) on NHWC bf16 (Unet.py:57) */
int pnnp_maxpool2x2_nhwc(const void* in, void* out, int n, int h, int w, int c, void* stream);

/* D2 — SynBase_Dataset.random_crop + data_aug (data_process/syn_datasets.py:100-107,162-173) for up to 64
 * crops of one packed frame (c x h x w fp32 -> n x c x patch x patch fp32): crop k is
 * frame[:, hs:hs+patch, ws:ws+patch] rotated by numpy.rot90(k = mode % 4) on (H, W) and W-flipped if mode >= 4.
 * The crop points / modes are host arrays (they come from the host RNG, init_random_crop_point :69-98). */
int pnnp_crop_aug(const float* frame, float* out, int c, int h, int w, int patch, int n,
                  const int* h_start_host, const int* w_start_host, const int* mode_host, void* stream);

/* White-balance jitter of Raw_Dataset.__getitem__ (data_process/syn_datasets.py:313-319; gains from random_gains,
 * data_process/unprocess.py:60-77), in place on n x c x h x w fp32 crops: every plane times rgb_gain (float32), then plane ch
 * times gain[ch] where kind[ch] = 1 (float32 product) or 2 (float64 product rounded to float32 once: what NumPy does when the
 * frame's white balance is np.float64); kind[ch] = 0 leaves the plane at the common gain.  kind / gain: host arrays, c entries. */
int pnnp_wb_gains(float* data, int n, int c, int h, int w, float rgb_gain, const int* kind_host, const double* gain_host,
                  void* stream);

/* HighBitRecovery.map (data_process/process.py:726-751): samples whose rounded DN value x is in [low, high) are re-drawn as
 * dist.ppf(cdf[x - low] + U * range[x - low]) (float64, rounded to float32), the sub-DN remainder is added back, then /span
 * (norm bit 0) or +bl; norm bit 1 = HighBitRecovery(float=False): the remainder is dropped.  dist = Tukey-lambda(lam) * scale + loc
 * (scipy's boxcox(u) - boxcox1p(-u), each term divided by lam), or N(loc, scale).  rand: caller-supplied U (replay) or NULL for
 * Philox4x32-10 draws keyed on (seed, offset, index0 + element); rand_out (optional) receives the U used. */
int pnnp_hbr_map(const float* in, float* out, size_t total, const double* cdf, const double* range, int low, int high,
                 int scale_in, int norm, float span, float bl, int dist_tukey, double lam, double loc, double scale,
                 const double* rand, uint64_t seed, uint64_t offset, uint64_t index0, double* rand_out, void* stream);

/* Overlapped tiling for tile-wise inference and its inverse — SynBase_Dataset.eval_crop / eval_merge
 * (data_process/syn_datasets.py:109-159): frame c x h x w fp32 <-> (h/l + 1)(w/l + 1) tiles of c x patch x patch,
 * l = patch - base, reflect padding base/2; the merge keeps each tile's interior l x l (later tiles win on overlaps). */
int pnnp_eval_crop(const float* frame, float* tiles, int c, int h, int w, int patch, int base, void* stream);
int pnnp_eval_merge(const float* tiles, float* frame, int c, int h, int w, int patch, int base, void* stream);

/* E1 / E2 — eval boundary on the device (trainer_SID.py:231-248; IlluminanceCorrect,
 * data_process/__init__.py:162-175; tensor2im + quality_assess, utils/visualization.py:9-31).
 * dn, hr: n x c x h x w fp32 (network output, clean target).  dn is scaled by `scale` (the ratio when
 * dst['ori'], else 1) and clamped to [0,1]; with brightness_correct the ELD gain <dn,hr>/<dn,dn> over
 * hr != 1 is applied; both images go through x255 + clip.  Writes per frame (3 + c) doubles into
 * `sums` (device): [num, den, sum of squared error, per-channel sums of the SSIM map over its
 * (h-6)(w-6) valid centres].  PSNR = 10 log10(255^2 c h w / sse); SSIM = mean_c(sum_c / ((h-6)(w-6))). */
int pnnp_eval_epilogue(const float* dn, const float* hr, int n, int c, int h, int w, float scale,
                       int brightness_correct, double* sums, void* stream);

/* ------------------------------------------------------------------------------------------
 * T1 — synthetic-pair training step (trainer_SID.py:93-101): L1 loss on pred.clamp(0,1)
 * (losses/base_loss.py:92-103), backward through the UNet, Adam (trainer_SID.py:44).
 * dgrad of a 3x3 conv = pnnp_conv2d_tc with transposed + flipped weights; wgrad = pnnp_wgrad_nhwc; the rest:
 * ------------------------------------------------------------------------------------------ */
/* loss_sum (device double) = sum |clamp(pred,0,1) - hr|;  gpred = d(mean L1)/d pred (NCHW fp32) */
int pnnp_l1_loss(const float* pred, const float* hr, float* gpred, size_t total, double* loss_sum, void* stream);
/* backward of the 1x1 head (conv10_1): gact = NHWC bf16 gradient w.r.t. the pre-activation of the layer feeding
 * it (act' applied), dW[co][cin] / db[co] / dbias_prev[cin] accumulated in fp32 */
int pnnp_head_bwd(const float* gpred, const void* act, const float* W, void* gact, float* dW, float* db,
                  float* dbias_prev, int n, int h, int w, int cin, int co, int act_kind, void* stream);
/* in place: g (NHWC bf16, w.r.t. activated output) *= act'(out); dbias[c] += sum over pixels (may be NULL) */
int pnnp_act_bwd_bias(void* g, const void* out, float* dbias, size_t pixels, int c, int act_kind, void* stream);
/* MaxPool2d(2) backward; gskip (optional) is added: gc = gskip + scatter(gp) (the U-Net skip connection); with act_kind != 0
 * the result is also multiplied by act'(cfull) (cfull is the activated output whose pooling this undoes) */
int pnnp_maxpool_bwd(const void* gp, const void* cfull, const void* gskip, void* gc, int n, int h, int w, int c,
                     int act_kind, void* stream);
/* the same with the bias gradient of the conv that produced cfull in the same pass: dbias[ch] += sum over pixels of gc (fp32
 * sums of the stored bf16 values — what pnnp_act_bwd_bias(gc, NULL, dbias, ..., act none) computes in a second pass over gc;
 * torch.autograd's conv bias gradient, /root/reference/trainer_SID.py:93-101 `loss.backward()`); c / 8 has to divide 256 */
int pnnp_maxpool_bwd_bias(const void* gp, const void* cfull, const void* gskip, void* gc, float* dbias, int n, int h, int w, int c,
                          int act_kind, void* stream);
/* Weight gradients straight from the NHWC bf16 activations (tcgen05, MN-major operands via TMA; no transposed copies):
 *   mode 0 (3x3 s1 p1 conv):      dw[ky*3+kx][ci_off + ci][co] += sum_p x[p + (ky-1, kx-1)][ci] * g[p][co]
 *   mode 1 (ConvTranspose2d 2x2): dw[a*2+b][ci_off + ci][co]   += sum_p x[p][ci] * g[2p + (a, b)][co]
 * g: [n][h|2h][w|2w][co_stride], x: [n][h][w][ci_stride]; dw: fp32 [taps][ci_total][co_pad] (zeroed by the caller).
 * co, ci: 16, 32, 64 or a multiple of 128. */
int pnnp_wgrad_nhwc(int mode, const void* g, int co, int co_stride, const void* x, int ci, int ci_stride, int n, int h, int w,
                    float* dw, int ci_off, int ci_total, int co_pad, void* stream);
int pnnp_wgrad_nhwc_pipeline_error(void);
/* Same update with the learning rate and the step count in device memory (state_dev[0] = lr, state_dev[1] = steps taken, as
 * floats; the call increments the step first), so a captured CUDA graph of the training step stays valid across steps. */
int pnnp_adam_step_dev(float* p, const float* g, float* m, float* v, size_t total, float* state_dev, float b1, float b2,
                       float eps, float gscale, void* stream);
/* Batched strided copy / cast: desc i copies dim[0] x dim[1] x dim[2] x dim[3] fp32 elements src[sum idx*sstride] ->
 * dst[sum idx*dstride] (fp32, or bf16 when dst_bf16).  Strides in elements, may be negative (src / dst point at index 0).
 * One launch for all descriptors (device array): weight packing for the tensor-core layouts and gradient re-layout. */
typedef struct pnnp_copy_desc {
    const void* src; void* dst;
    int dst_bf16; int dim[4];
    long long sstride[4]; long long dstride[4];
} pnnp_copy_desc;
int pnnp_strided_copy_batch(const pnnp_copy_desc* descs_dev, int n_desc, int blocks_per_desc, void* stream);
/* torch.optim.Adam step (no weight decay) over flat fp32 buffers; g is multiplied by gscale first */
int pnnp_adam_step(float* p, const float* g, float* m, float* v, size_t total, float lr, float b1, float b2, float eps,
                   int step, float gscale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PNNP_B200_H */
