// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by
// TMA, for the UNet / ResUnet denoiser forward (archs/Unet.py:54-99, archs/ResUnet.py:46-88).
//
// One persistent, warp-specialised kernel covers every GEMM-shaped layer of both networks:
//   mode CONV3  3x3, pad 1, stride 1   (conv*_1/_2, ResidualBlock convs)     taps = 9
//   mode CONV1  1x1                     (conv10_1, ResidualBlock short_cut)    taps = 1
//   mode CONVT  ConvTranspose2d(2, s2)  (upv6..upv9): 4 GEMMs, one per (a,b), pixel-shuffle store
// Activations are NHWC bf16 (channels padded to a multiple of 16), accumulation fp32 in TMEM.
//
// GEMM view.  M = 128 output pixels (an 8 x 16 spatial tile), N = UMMA_N output channels,
// K = taps x Cin.  The A operand of tap (dy,dx) is the input tile shifted by (dy-1, dx-1); a 4-D
// TMA box (Kc channels x 16 x (8+2) x 1) lands in shared memory as 160 rows of Kc*2 bytes in the
// canonical K-major swizzled layout, with out-of-bounds rows/columns zero-filled by the TMA unit
// (= the convolution's zero padding).  One box per dx serves the three dy taps: the UMMA
// descriptor start address is advanced by dy*16 rows (a whole number of swizzle atoms).  The
// channel concat of the decoder (torch.cat([up, skip], 1), Unet.py:72) is never materialised: the
// K loop simply walks two tensor maps.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected lane), warps 2-5 = epilogue (TMEM -> registers -> bias/activation/residual -> global).
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), TMEM accumulator double buffer
// full/empty mbarriers (MMA <-> epilogue).
#ifndef PNNP_HOST_EMUL
#include <cuda.h>
#include <cuda_bf16.h>
#endif
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "abi_common.h"
#include "tc_common.cuh"
#include "../../include/pnnp_b200.h"

// kernel launch of the (taps per stage, K16 slices, epilogue, variant) instantiation; tests/emul/ compiles this file for the host and
// runs the launch on its SIMT emulator + tensor-core model instead (test infrastructure only)
#ifdef PNNP_HOST_EMUL
#define PNNP_CONV_KLAUNCH(T, K, E, V) emul_launch_1d(grid, 64 + 128 * groups, [&]() { conv_gemm_tc_kernel<T, K, E, V>(tmA0, tmA1, tmB, p); })
#else
#define PNNP_CONV_KLAUNCH(T, K, E, V) conv_gemm_tc_kernel<T, K, E, V><<<grid, 64 + 128 * groups, smem, st>>>(tmA0, tmA1, tmB, p)
#endif

namespace pnnp {

constexpr int kTileW = 16, kTileH = 8, kTileM = 128;
constexpr int kMaxGroups = 4;            // epilogue groups == TMEM accumulator buffers (2 when N > 128, else 4)
constexpr int kConvThreadsMax = 64 + 128 * kMaxGroups;   // warp 0 TMA, warp 1 MMA, then 4 epilogue warps per group
constexpr int kMaxStages = 8;

enum { MODE_CONV3 = 0, MODE_CONV1 = 1, MODE_CONVT = 2, MODE_CONV3S2 = 3, MODE_CONV3X = 4, MODE_CONV2S2 = 5, MODE_CONV3B = 6 };
// MODE_CONV3B ("one box"): all NINE taps of a 3x3 conv from ONE haloed TMA box.  The tile is 8 pixels wide x 16 tall, so a tile row
// is exactly one 8-row core-matrix group of the K-major operand and the stride between groups (SBO) is the box's row pitch, 10 pixels;
// tap (dy, dx) is the start-address offset (dy * 10 + dx) pixel rows.  tcgen05.mma applies the shared-memory swizzle to the ABSOLUTE
// address (tools/ubench_umma_offset.cu, measured on a B200 in r02: any start offset in whole rows and any SBO read exactly the rows
// of the linear-address model, base-offset field 0), which is also how TMA wrote them.  Against the x-shift-in-N mode: the
// accumulator has Cout columns instead of 3 Cout (a third of the TMEM drain, which at 64 B per clock was 768 of a tile's ~1 200
// cycles), no cross-lane combine, no halo lanes; against the per-tap mode: one box and one hand-shake per tile instead of three.
constexpr int kTileWB = 8, kTileHB = 16, kBoxWB = kTileWB + 2, kBoxHB = kTileHB + 2;
constexpr int kTileWX = 14;             // output columns per tile in MODE_CONV3X (16 partial-sum columns, 1 halo each side)
enum { OUT_NHWC_BF16 = 0, OUT_NCHW_F32 = 1 };
enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2 };

struct ConvParams {
    int mode, n_img, H, W;          // H, W: tiled spatial grid = output size (CONV3/CONV1/CONV3S2) or input size (CONVT: output 2H x 2W)
    int tiles_x, tiles_y, n_tiles;
    int umma_n;                     // columns per accumulator / MMA N
    int nsrc, cin0, cin1;           // channels of the two K sources (multiples of kc); cin1 = 0 if single
    int kc, swz;                    // channels per K chunk (16/32/64) and swizzle span in bytes (= 2*kc)
    int cout;                       // real output channels written per pixel
    int cout_stride;                // channel stride of the NHWC output (>= cout)
    int act, out_mode;
    int stages, stage_bytes, a_bytes, b_tap_stride;
    int b_resident;                 // all taps x K chunks of the weights stay in smem for the whole kernel
    int b_res_bytes;
    int tmem_cols;
    int groups;                     // epilogue groups / accumulator buffers in flight
    const float* bias;              // [cout] or null
    void* out;
    const __nv_bfloat16* resid;     // NHWC bf16 residual with the output's geometry, or null
    const float* resid_nchw;        // NCHW fp32 residual for OUT_NCHW_F32 / the fused head, or null
    const __nv_bfloat16* mask;      // NHWC bf16 activation with the output's geometry: out *= (mask > 0 ? 1 : mask_slope), or null
    float mask_slope;               //   (LeakyReLU' / ReLU' of the layer whose data gradient this conv computes)
    __nv_bfloat16* pool_out;        // optional fused 2x2 max-pool output (NHWC bf16, H/2 x W/2)
    const float* head_w;            // optional fused 1x1 head: [head_cout][cout] fp32 weights
    const float* head_b;            // [head_cout]
    float* head_out;                // NCHW fp32, head_cout <= 4 planes
    int head_cout;
    int dbg;                        // timing experiments only: 1 skip stores, 2 skip MMAs, 4 skip A loads, 8 skip epilogue math
    int* err;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

// Tile walk.  A CTA owns a CONTIGUOUS range of tiles (order: n_tile fastest, then x, y, image), so neighbouring tiles' halos
// meet in L2 while they are hot and the (n_tile, x, y, image) coordinates advance by increment-and-carry: the per-tile index
// arithmetic (three run-time divisions per tile in each of the three warp roles) was, for the small-K full-resolution
// layers, most of a single warp's serial instruction stream per tile — ~800 cycles per tile with loads, MMAs and epilogue all
// switched off (r01 experiment PNNP_CONV_DBG=14).
struct TileIter {
    // Order (r02): images, then BANDS of kBand tile rows, inside a band tile columns, inside a column the band's rows, n_tile
    // fastest.  Vertically adjacent tiles are then consecutive in a CTA's range, so the 2 halo rows a tile shares with its
    // neighbour (25 % of a 10-row box) are still in L2 when they are fetched again, and horizontal neighbours are only kBand tiles
    // apart.  With the row-major walk of round 1 a CTA came back to the next tile row ~250 MB of traffic later — past the 126 MB
    // L2 — and the full-resolution layers read 23-25 % more from DRAM than their inputs hold (ncu, r02).
    static constexpr int kBand = 4;
    int t, t_end, n_tile, tx, ty, img, band0, rows, ry;
    __device__ __forceinline__ void init(int total_tiles, int n_tiles, int tiles_x, int tiles_y) {
        const int per = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
        t = min(total_tiles, (int)blockIdx.x * per);
        t_end = min(total_tiles, t + per);
        int r = t;
        n_tile = r % n_tiles; r /= n_tiles;
        const int per_img = tiles_x * tiles_y;
        img = r / per_img;
        int q = r - img * per_img;
        const int band = q / (kBand * tiles_x);
        band0 = band * kBand;
        rows = min(kBand, tiles_y - band0);
        q -= band * kBand * tiles_x;
        tx = q / rows;
        ry = q - tx * rows;
        ty = band0 + ry;
    }
    __device__ __forceinline__ bool valid() const { return t < t_end; }
    __device__ __forceinline__ void next(int n_tiles, int tiles_x, int tiles_y) {
        ++t;
        if (++n_tile == n_tiles) {
            n_tile = 0;
            if (++ry == rows) {
                ry = 0;
                if (++tx == tiles_x) {
                    tx = 0;
                    band0 += kBand;
                    if (band0 >= tiles_y) { band0 = 0; ++img; }
                    rows = min(kBand, tiles_y - band0);
                }
            }
            ty = band0 + ry;
        }
    }
};

// All MMAs of one pipeline stage: TPS taps x K16S K-slices, fully unrolled; descriptors advance by adding
// to the 14-bit start-address field (no carry out of the field: shared memory is < 256 KB).
template <int TPS, int K16S, bool F32 = false>
__device__ __forceinline__ void issue_stage_mmas(uint32_t d_tmem, uint64_t adesc0, uint64_t bdesc0, uint32_t a_tap_stride,
                                                 uint32_t b_tap_stride, uint32_t idesc, bool accumulate_first) {
    if constexpr (TPS == 9) {
        // MODE_CONV3B: a_tap_stride = one pixel row of the box in descriptor units; tap (dy, dx) starts (dy * kBoxWB + dx) rows in.
        // Taps in the per-tap mode's order (dx outer, dy inner, K slices innermost), so the fp32 accumulators — not only the rounded
        // outputs — are bit-identical to MODE_CONV3's.
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int dx = t / 3, dy = t % 3;
#pragma unroll
            for (int k = 0; k < K16S; ++k) {
                // both offsets are compile-time constants (a pixel row of the box is 32 K16S bytes; the launcher lays the weights of
                // a tap out 64 rows apart whatever Cout is), so a descriptor is ONE 64-bit add with an immediate: the MMA warp's
                // instruction stream is what bounds this mode (r02 capture: 186 instructions per tile, a third of them zero-extending
                // and adding run-time tap strides)
                constexpr uint32_t kARow = (32u * K16S) >> 4, kBTap = (64u * 32u * K16S) >> 4;
                const uint64_t ad = adesc0 + (uint64_t)((dy * kBoxWB + dx) * kARow + k * 2);
                const uint64_t bd = bdesc0 + (uint64_t)((dy * 3 + dx) * kBTap + k * 2);
                tc_mma_bf16(d_tmem, ad, bd, idesc, (accumulate_first || t != 0 || k != 0) ? 1u : 0u);
            }
        }
        return;
    }
#pragma unroll
    for (int dy = 0; dy < TPS; ++dy) {
#pragma unroll
        for (int k = 0; k < K16S; ++k) {
            const uint64_t ad = adesc0 + (uint64_t)(dy * a_tap_stride + k * 2);     // k * 32 B
            const uint64_t bd = bdesc0 + (uint64_t)(dy * b_tap_stride + k * 2);
            if constexpr (F32) tc_mma_tf32(d_tmem, ad, bd, idesc, (accumulate_first || dy != 0 || k != 0) ? 1u : 0u);   // 32 bytes = 8 fp32 per slice
            else tc_mma_bf16(d_tmem, ad, bd, idesc, (accumulate_first || dy != 0 || k != 0) ? 1u : 0u);
        }
    }
}

// ------------------------------------------------------------------------------------------ kernel
// EPI: compile-time specialisation of the epilogue.  -1 = generic (every feature decided at run time).  >= 0 = the hot NHWC-output
// 3x3 layers (no residual, no transposed conv), bit 0 = x-shift-in-N mode, bit 1 = fused 2x2 max-pool, bit 2 = fused 1x1 head:
// the narrow full-resolution layers are bound by the epilogue's instruction count, and most of it was run-time feature tests.
constexpr int EPI_GENERIC = -1, EPI_X = 1, EPI_POOL = 2, EPI_HEAD = 4, EPI_MASK = 8;   // bit 3: activation-derivative mask (training dgrad)
constexpr int EPI_RESID = 32;          // bit 5: NHWC bf16 residual added after the activation (the second conv of a ResUnet residual block; r02: those layers ran the generic epilogue, 217 us against 135 for the same conv without the residual)
constexpr int EPI_CONVT = 16;           // bit 4: ConvTranspose2d pixel-shuffle store (bf16 NHWC, bias only) — default for > 64 input channels since r02, see conv_layer_launch
// SUP: super-tile.  One pipeline stage carries a 16-row box (8 + 8 + 2 halo rows, ONE TMA load) and feeds TWO M = 128 tiles —
// rows 0-7 into accumulator buffer a, rows 8-15 (A descriptor start + 128 pixel rows) into buffer a + 1 — so the producer <-> MMA
// hand-shake, which bounds the small-K full-resolution layers, is paid once per 256 pixels, and a tile's halo
// rows are 2 in 18 instead of 2 in 10.  Epilogue group g drains buffer g = half (g & 1) of every (groups / 2)-th super-tile.
// VAR bit 0 = SUP (above); bit 1 = PDL: the kernel is launched with programmatic stream serialization and waits for the previous
// grid of the stream before its first global-memory access (see the top of the body).  Both are compile-time so that the
// default instantiations (VAR = 0) keep exactly the machine code that was measured in round 1.
// VAR bit 2 = X2: the epilogue's fp32 arithmetic (x-mode partial-sum combine, bias + activation, fused 1x1 head) in packed pairs —
// the same IEEE additions / FMAs, half the instructions (default since r02; PNNP_CONV_F32X2=0 switches it off).
template <int TPS, int K16S, int EPI, int VAR = 0>
__global__ void __launch_bounds__(kConvThreadsMax, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    constexpr int SUP = VAR & 1;
    constexpr bool kPdl = (VAR & 2) != 0;
    constexpr bool kX2 = (VAR & 4) != 0;
    // VAR bit 3 = F32: the fp32-storage variant (north_star: "within 1e-3 in fp32, bf16 variant reported separately") — activations and
    // weights are fp32 in HBM and shared memory, the MMAs are kind::tf32 (K = 8 per 32-byte slice, so the same slice / swizzle /
    // descriptor arithmetic applies with kc = span / 4 channels per chunk), the epilogue stores fp32 NHWC.  Generic epilogue only.
    constexpr bool kF32 = (VAR & 8) != 0;
    // Variant instantiations (VAR != 0; the defaults since r02) with a specialised epilogue are never launched with debug switches, so their mode (bits 0 / 4 of
    // EPI), the debug word and the swizzle span (32 bytes per K16 slice) are compile-time constants: the serial loops of the
    // producer and MMA warps — whose instruction count IS the tile rate of the small-K layers — lose their run-time selects.
    constexpr bool kCt = VAR != 0 && EPI >= 0;       // every variant instantiation with a specialised epilogue (the launcher never
    constexpr bool kB = TPS == 9;                    // MODE_CONV3B: 8 x 16 tiles, one box, nine taps by descriptor offset
    constexpr int kTW = kB ? kTileWB : kTileW, kTH = kB ? kTileHB : kTileH;
    constexpr int kCtMode = kB ? MODE_CONV3B : ((EPI >= 0 && (EPI & EPI_CONVT)) ? MODE_CONVT : ((EPI >= 0 && (EPI & EPI_X)) ? MODE_CONV3X : MODE_CONV3));   // pairs those with debug switches)
#define PNNP_MODE_K (kCt ? kCtMode : p.mode)        /* expressions, not locals: the default instantiations must compile exactly as before */
#define PNNP_DBG_K (kCt ? 0 : p.dbg)
#define PNNP_SWZ_K (kCt ? 32 * K16S : p.swz)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x stage_bytes] [barriers] [tmem slot] [bias]
    // 1024-byte alignment by POINTER arithmetic on the __shared__ array (not through an integer cast): the compiler keeps the
    // shared address space of everything carved from it, so plain loads / stores of these pointers are LDS / STS instead of
    // generic LD / ST (which go through the global-memory instruction queue: stall reason lg_throttle in the r02 capture)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_bres = smem + (size_t)p.stages * p.stage_bytes;          // resident weights (1024-aligned), may be empty
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bres + p.b_res_bytes);
    uint64_t* full_bar = bars;                       // [stages]
    uint64_t* empty_bar = bars + kMaxStages;         // [stages]
    uint64_t* tfull_bar = bars + 2 * kMaxStages;                    // [kMaxGroups]
    uint64_t* tempty_bar = bars + 2 * kMaxStages + kMaxGroups;      // [kMaxGroups]
    uint64_t* bres_bar = bars + 2 * kMaxStages + 2 * kMaxGroups;    // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kMaxGroups + 2);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_head_w = s_bias + p.cout;            // [cout][4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int taps_per_stage = TPS;
    const int dx_count = PNNP_MODE_K == MODE_CONV3 ? 3 : (PNNP_MODE_K == MODE_CONV3S2 ? 9 : (PNNP_MODE_K == MODE_CONV2S2 ? 4 : 1));   // pipeline stages per K chunk
    const int chunks0 = p.cin0 / p.kc, chunks1 = p.nsrc > 1 ? p.cin1 / p.kc : 0;
    const int ksteps = (chunks0 + chunks1) * dx_count;
    const int total_tiles = p.n_img * p.tiles_y * p.tiles_x * p.n_tiles;
    const uint32_t stage_tx = (uint32_t)(p.a_bytes + (p.b_resident ? 0 : taps_per_stage * p.umma_n * PNNP_SWZ_K));
    const int taps_total = (PNNP_MODE_K == MODE_CONV3 || PNNP_MODE_K == MODE_CONV3B) ? 9 : (PNNP_MODE_K == MODE_CONV3S2 ? 9 : (PNNP_MODE_K == MODE_CONV3X ? 3 : (PNNP_MODE_K == MODE_CONV2S2 ? 4 : 1)));

    if constexpr (kPdl) {
        // Programmatic dependent launch (default since r02; PNNP_CONV_PDL=0 switches it off): this grid may have been scheduled while the previous kernel of
        // the stream was still draining.  Nothing above touches global memory; every thread waits here for the previous grid to
        // complete and flush, then lets the next conv layer's CTAs be scheduled as ours retire.
        pdl_wait_then_release();
    }
    for (int i = threadIdx.x; i < p.cout; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (p.head_out)
        for (int i = threadIdx.x; i < 4 * p.cout; i += blockDim.x)      // [channel][4 outputs], zero beyond head_cout
            s_head_w[i] = (i & 3) < p.head_cout ? p.head_w[(i & 3) * p.cout + (i >> 2)] : 0.f;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmA0);
        prefetch_tensormap(&tmA1);
        prefetch_tensormap(&tmB);
        for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int a = 0; a < p.groups; ++a) { mbar_init(smem_u32(&tfull_bar[a]), 1); mbar_init(smem_u32(&tempty_bar[a]), 4); }
        mbar_init(smem_u32(bres_bar), 1);
        fence_mbarrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Producer and MMA warps run their loops warp-uniformly (all 32 lanes take the same branches and
    // hold the same values, so addresses/descriptors stay in uniform registers); only the TMA / MMA /
    // commit instructions themselves are predicated on one elected lane.
    if (warp == 0) {
        // ============================== TMA producer ==============================
        // Everything the loops need is copied into registers first: the asm statements carry "memory" clobbers, so every p.field
        // used inside a loop would otherwise be re-read from the parameter bank per tile, and for the small-K full-resolution
        // layers (one pipeline stage per tile) this single warp's serial instruction stream IS the tile rate.
        const int mode = PNNP_MODE_K, n_tiles = p.n_tiles, tiles_x = p.tiles_x, tiles_y = p.tiles_y, umma_n = p.umma_n, kc = kCt ? 16 * K16S : p.kc;
        const int cin0 = p.cin0, stages = p.stages, stage_bytes = p.stage_bytes, a_bytes = p.a_bytes, b_tap_bytes = p.b_tap_stride;
        const int dbg = PNNP_DBG_K;
        const bool b_res = p.b_resident != 0;
        int* const err = p.err;
        const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        const int halo = (mode == MODE_CONV3 || mode == MODE_CONV3B) ? 1 : 0;
        const int yhalo = (mode == MODE_CONV3 || mode == MODE_CONV3X || mode == MODE_CONV3B) ? 1 : 0;
        const int tile_w = mode == MODE_CONV3X ? kTileWX : kTW, x_first = mode == MODE_CONV3X ? -1 : 0;
        const bool plain = mode != MODE_CONV2S2 && mode != MODE_CONV3S2;
        if (b_res && elect_one()) {
            // weights for every (K chunk, tap) once per CTA: layout [chunk][tap][umma_n rows x swz bytes]
            const uint32_t bb = smem_u32(bres_bar);
            mbar_expect_tx(bb, (uint32_t)((chunks0 + chunks1) * taps_total * umma_n * PNNP_SWZ_K));
            for (int ch = 0; ch < chunks0 + chunks1; ++ch)
                for (int tap = 0; tap < taps_total; ++tap)
                    tma_load_3d(smem_u32(smem_bres) + (ch * taps_total + tap) * b_tap_bytes, &tmB, bb, ch * kc, 0, tap);
        }
        __syncwarp();
        TileIter ti;
        ti.init(total_tiles, n_tiles, tiles_x, tiles_y);
        const int first_tile = ti.t;
        uint32_t stage = 0, phase = 0, sa = smem0, fb = full0, eb = empty0;
        for (; ti.valid(); ti.next(n_tiles, tiles_x, tiles_y)) {
            const int img = ti.img;
            const int x0 = ti.tx * tile_w + x_first, y0 = ti.ty * (SUP ? 2 * kTileH : kTH);
            const int n_off = ti.n_tile * umma_n;
            const bool skip_loads = (dbg & 4) && ti.t != first_tile;
            int chunk = 0, dx = 0;
            for (int ks = 0; ks < ksteps; ++ks) {
                const int src = chunk >= chunks0 ? 1 : 0;
                const int cc = src ? chunk - chunks0 : chunk;
                mbar_wait(eb, phase ^ 1, err, 101);
                if (skip_loads) { if (elect_one()) mbar_arrive(fb); }
                else if (elect_one()) {
                    mbar_expect_tx(fb, stage_tx);
                    if (plain)
                        tma_load_4d(sa, src ? &tmA1 : &tmA0, fb, cc * kc, x0 + dx - halo, y0 - yhalo, img);
                    else if (mode == MODE_CONV2S2)   // 2x2 stride 2, no padding (ConvTranspose2d dgrad): tap = a*2+b
                        tma_load_4d(sa, &tmA0, fb, cc * kc, 2 * x0 + (dx & 1), 2 * y0 + (dx >> 1), img);
                    else                             // stride 2: one box per tap, TMA element stride 2 in x and y
                        tma_load_4d(sa, &tmA0, fb, cc * kc, 2 * x0 + (dx % 3) - 1, 2 * y0 + (dx / 3) - 1, img);
                    if (!b_res) {
                        const int cin_off = (src ? cin0 : 0) + cc * kc;
                        for (int dy = 0; dy < taps_per_stage; ++dy) {
                            const int tap = mode == MODE_CONV3 ? dy * 3 + dx
                                          : ((mode == MODE_CONV3S2 || mode == MODE_CONV2S2) ? dx : ((mode == MODE_CONV3X || mode == MODE_CONV3B) ? dy : 0));
                            tma_load_3d(sa + a_bytes + dy * b_tap_bytes, &tmB, fb, cin_off, n_off, tap);
                        }
                    }
                }
                __syncwarp();
                if (++dx == dx_count) { dx = 0; ++chunk; }
                sa += stage_bytes; fb += 8; eb += 8;
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; sa = smem0; fb = full0; eb = empty0; }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
        const int mode = PNNP_MODE_K, umma_n = p.umma_n, stages = p.stages, stage_bytes = p.stage_bytes, a_bytes = p.a_bytes, dbg = PNNP_DBG_K;
        const int groups = p.groups;
        const bool b_res = p.b_resident != 0;
        int* const err = p.err;
        const uint32_t idesc = (1u << 4) | ((kF32 ? 2u : 1u) << 7) | ((kF32 ? 2u : 1u) << 10) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);   // A/B format: 1 = bf16, 2 = tf32
        const uint64_t dhi = umma_desc_hi(PNNP_SWZ_K);
        // MODE_CONV3B: the A operand's 8-row groups are kBoxWB pixel rows apart (SBO), and a tap advances by whole pixel rows
        const uint64_t dhi_a = kB ? ((dhi & ~(0x3FFFull << 32)) | ((uint64_t)((kBoxWB * PNNP_SWZ_K) >> 4) << 32)) : dhi;
        const uint32_t a_tap_stride = kB ? (uint32_t)PNNP_SWZ_K >> 4 : (uint32_t)(kTileW * PNNP_SWZ_K) >> 4;      // descriptor units (16 B)
        const uint32_t b_tap_stride = (uint32_t)p.b_tap_stride >> 4;
        const uint32_t b_tap_bytes = (uint32_t)p.b_tap_stride;
        const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar), bres0 = smem_u32(smem_bres);
        // resident layout [chunk][tap]: CONV3 stage dx uses taps dy*3+dx (stride 3 taps); CONV3X: dy taps are consecutive
        const uint32_t b_stride_eff = (b_res && mode == MODE_CONV3) ? 3 * b_tap_stride : b_tap_stride;
        const uint32_t a_half = (uint32_t)(kTileH * kTileW * PNNP_SWZ_K) >> 4;   // SUP: second tile's A rows start 8 x 16 pixel rows further
        if (b_res) mbar_wait(smem_u32(bres_bar), 0, err, 105);
        const int dxc = mode == MODE_CONV3 ? 3 : (mode == MODE_CONV3S2 ? 9 : (mode == MODE_CONV2S2 ? 4 : 1));
        const int per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
        const int t_begin = min(total_tiles, (int)blockIdx.x * per_cta), t_stop = min(total_tiles, t_begin + per_cta);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, sa = smem0, fb = full0, eb = empty0;
        for (int t = t_begin; t < t_stop; ++t) {
            if (!(dbg & 32)) mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1, err, 102);
            if (SUP && !(dbg & 32)) mbar_wait(tempty0 + 8 * (acc + 1), acc_phase ^ 1, err, 106);
            if (!(dbg & 64)) tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * (uint32_t)umma_n;
            uint32_t sb_res = bres0;                                       // resident weights of (chunk, dx): advances by one tap
            int dx = 0;
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(fb, phase, err, 103);
                if (!(dbg & 64)) tc_fence_after();
                const uint64_t adesc0 = dhi_a | (uint64_t)((sa >> 4) & 0x3FFFu);
                const uint32_t sb = b_res ? sb_res : sa + (uint32_t)a_bytes;
                const uint64_t bdesc0 = dhi | (uint64_t)((sb >> 4) & 0x3FFFu);
                // (chunk * taps_total + dx) * tap bytes: +1 tap per stage, +taps_total per chunk
                if (++dx == dxc) { dx = 0; sb_res += (uint32_t)(taps_total - dxc + 1) * b_tap_bytes; } else sb_res += b_tap_bytes;
                if (elect_one()) {
                    if (!(dbg & 2)) issue_stage_mmas<TPS, K16S, kF32>(d_tmem, adesc0, bdesc0, a_tap_stride, b_stride_eff, idesc, ks != 0);
                    if (SUP && !(dbg & 2))
                        issue_stage_mmas<TPS, K16S, kF32>(d_tmem + (uint32_t)umma_n, adesc0 + a_half, bdesc0, a_tap_stride, b_stride_eff, idesc, ks != 0);
                    if (dbg & 16) mbar_arrive(eb);
                    else tc_commit(eb);                                    // frees the smem slot when these MMAs retire
                }
                __syncwarp();
                sa += (uint32_t)stage_bytes; fb += 8; eb += 8;
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; sa = smem0; fb = full0; eb = empty0; }
            }
            if (elect_one() && !(dbg & 32)) {                              // accumulator(s) complete -> epilogue
                if (dbg & 16) mbar_arrive(tfull0 + 8 * acc); else tc_commit(tfull0 + 8 * acc);
                if (SUP) { if (dbg & 16) mbar_arrive(tfull0 + 8 * (acc + 1)); else tc_commit(tfull0 + 8 * (acc + 1)); }
            }
            __syncwarp();
            acc += SUP ? 2 : 1;
            if (acc == (uint32_t)groups) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ============================== epilogue (warps 2..5) ==============================
        // Epilogue groups of four warps: group g drains accumulator buffer g, i.e. every groups-th tile of this
        // CTA, so several tiles' epilogues (latency-bound: TMEM load -> math -> shuffle -> store chains) overlap
        // each other and the following tiles' MMAs.
        const int group = (warp - 2) >> 2;
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
        const int m = quad * 32 + lane;                    // accumulator row == pixel within the tile
        const int ty_in = m / kTW, tx_in = m % kTW;
        const uint32_t acc = (uint32_t)group;
        uint32_t acc_phase = 0;
        constexpr bool kSpec = EPI >= 0;
        const bool xmode = kSpec ? (EPI & EPI_X) != 0 : p.mode == MODE_CONV3X;
        const bool is_convt = kSpec ? (EPI & EPI_CONVT) != 0 : p.mode == MODE_CONVT;
        const bool out_nhwc = kSpec ? true : p.out_mode == OUT_NHWC_BF16;
        const bool has_resid = kSpec ? (EPI & EPI_RESID) != 0 : p.resid != nullptr;
        const bool has_mask = kSpec ? (EPI & EPI_MASK) != 0 : p.mask != nullptr;
        const bool has_pool = kSpec ? (EPI & EPI_POOL) != 0 : p.pool_out != nullptr;
        const bool has_head = kSpec ? (EPI & EPI_HEAD) != 0 : p.head_out != nullptr;
        // activation as max(v, v * slope + 0): LeakyReLU 0.2 / ReLU (slope 0; the +0 turns -0 into +0) / identity (slope 1)
        const float slope = p.act == ACT_LEAKY ? 0.2f : (p.act == ACT_RELU ? 0.f : 1.f);
        // Every parameter the tile loop reads, copied out of the parameter bank once: the asm statements carry "memory" clobbers, so each
        // p.field inside the loop was re-read per tile (the r02 capture of conv9_2 + head shows LDCU -> compare -> branch chains on the
        // critical path of every tile: 15 % of the epilogue warps' stall samples were per-tile set-up)
        const int eH = p.H, eW = p.W, e_cout = p.cout, e_umma_n = p.umma_n, e_n_tiles = p.n_tiles, e_tiles_x = p.tiles_x, e_tiles_y = p.tiles_y;
        const int e_cout_stride = p.cout_stride, e_groups = p.groups;
        int* const e_err = p.err;
        void* const e_out = p.out;
        void* const e_pool_out = p.pool_out;
        const __nv_bfloat16* const e_resid = p.resid;
        const __nv_bfloat16* const e_mask = p.mask;
        const float e_mask_slope = p.mask_slope;
        const int chunks16 = (xmode ? e_cout : e_umma_n) / 16;
        // fused head: its four biases once per thread (they were four global loads per pixel and tile — 13 % of the stall
        // samples of conv9_2 + head in the r02 capture, each addition waiting for its own load)
        float head_bias[4] = {0.f, 0.f, 0.f, 0.f};
        if (has_head) {
#pragma unroll
            for (int o = 0; o < 4; ++o) head_bias[o] = o < p.head_cout ? p.head_b[o] : 0.f;
        }
        const int head_cout = p.head_cout;
        float* const head_out = p.head_out;
        const float* const head_resid = p.resid_nchw;
        TileIter ti;
        ti.init(total_tiles, e_n_tiles, e_tiles_x, e_tiles_y);
        if (PNNP_DBG_K & 32) ti.t = ti.t_end;
        const int slot = SUP ? group >> 1 : group, n_slots = SUP ? e_groups >> 1 : e_groups;     // SUP: two groups share a super-tile
        const int y_in = SUP ? (group & 1) * kTileH + ty_in : ty_in, tile_rows = SUP ? 2 * kTileH : kTH;
        for (int g = 0; g < slot && ti.valid(); ++g) ti.next(e_n_tiles, e_tiles_x, e_tiles_y);   // slot s: every n_slots-th tile
        for (; ti.valid(); ) {
            const int n_tile = ti.n_tile, tx = ti.tx, ty = ti.ty, img = ti.img;
            for (int g = 0; g < n_slots && ti.valid(); ++g) ti.next(e_n_tiles, e_tiles_x, e_tiles_y);
            const int x = xmode ? tx * kTileWX - 1 + tx_in : tx * kTW + tx_in, y = ty * tile_rows + y_in;
            const bool valid = x < eW && y < eH && (!xmode || (tx_in >= 1 && tx_in <= kTileWX));
            mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, e_err, 104);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * (uint32_t)e_umma_n + ((uint32_t)(quad * 32) << 16);
            if (PNNP_DBG_K & 8) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc])); acc_phase ^= 1; continue; }
            const int col_tile0 = n_tile * e_umma_n;     // first GEMM column of this tile
            const size_t pix_in = ((size_t)img * eH + y) * (size_t)eW + x;
            // per-tile addresses hoisted out of the chunk loop (the asm statements' memory clobbers make the compiler re-read the
            // parameter bank and redo the 64-bit products per 16-channel chunk otherwise)
            const size_t pool_off = has_pool ? (((size_t)img * (eH >> 1) + (y >> 1)) * (size_t)(eW >> 1) + (x >> 1)) * (size_t)e_cout_stride : 0;
            const size_t out_off = pix_in * (size_t)e_cout_stride;
            float head[4] = {0.f, 0.f, 0.f, 0.f};
            // specialised ConvTranspose2d epilogue: (tap, channel block) of chunk j kept as a running pair — one division per tile
            constexpr bool kCtT = EPI >= 0 && (EPI & EPI_CONVT) != 0;
            int t_cpt = 1, t_tap = 0, t_rem = 0;
            if constexpr (kCtT) { t_cpt = e_cout >> 4; t_tap = (col_tile0 >> 4) / t_cpt; t_rem = (col_tile0 >> 4) - t_tap * t_cpt; }
            for (int j = 0; j < chunks16; ++j) {
                uint32_t v[16];
                tc_ld16(taddr + j * 16, v);
                if (xmode) {
                    // columns are (dx, co): out(j) = P[j-1, dx=0] + P[j, dx=1] + P[j+1, dx=2] along the 16-lane tile row
                    uint32_t v0[16], v2[16];
                    tc_ld16(taddr + e_cout + j * 16, v0);          // dx = 1 block (own column)
                    tc_ld16(taddr + 2 * e_cout + j * 16, v2);      // dx = 2 block
                    tc_ld_wait();
                    if constexpr (kX2) {
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const float l0 = __shfl_up_sync(0xffffffffu, __uint_as_float(v[i]), 1, 16);
                            const float l1 = __shfl_up_sync(0xffffffffu, __uint_as_float(v[i + 1]), 1, 16);
                            const float r0 = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[i]), 1, 16);
                            const float r1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[i + 1]), 1, 16);
                            float s0, s1;          // (left + own) + right, as in the scalar form
                            f2_unpack(f2_add(f2_add(f2_pack(l0, l1), f2_pack(__uint_as_float(v0[i]), __uint_as_float(v0[i + 1]))), f2_pack(r0, r1)), s0, s1);
                            v[i] = __float_as_uint(s0); v[i + 1] = __float_as_uint(s1);
                        }
                    } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v[i]), 1, 16);
                        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[i]), 1, 16);
                        v[i] = __float_as_uint(left + __uint_as_float(v0[i]) + right);
                    }
                    }
                } else {
                    tc_ld_wait();
                }
                int c0 = col_tile0 + j * 16;              // output channel of v[0]
                size_t opix = pix_in;
                if constexpr (kCtT) {
                    const int tap = t_tap;
                    c0 = t_rem << 4;
                    if (++t_rem == t_cpt) { t_rem = 0; ++t_tap; }
                    opix = ((size_t)img * (2 * eH) + (2 * y + (tap >> 1))) * (size_t)(2 * eW) + (2 * x + (tap & 1));
                } else if (is_convt) {                    // GEMM column = (a*2+b)*cout + co  ->  pixel (2y+a, 2x+b)
                    const int tap = c0 / e_cout;
                    c0 -= tap * e_cout;
                    opix = ((size_t)img * (2 * eH) + (2 * y + (tap >> 1))) * (size_t)(2 * eW) + (2 * x + (tap & 1));
                }
                if (c0 >= e_cout) continue;               // zero-padded weight rows (warp-uniform)
                float f[16];
                if constexpr (kCtT) {                     // bias only: the launcher takes this epilogue for act == NONE (identity) alone
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * i4);
                        f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1])), f2_pack(b4.x, b4.y)), f[4 * i4], f[4 * i4 + 1]);
                        f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3])), f2_pack(b4.z, b4.w)), f[4 * i4 + 2], f[4 * i4 + 3]);
                    }
                } else if (kX2 && out_nhwc) {
                    const uint64_t slope2 = f2_pack(slope, slope), zero2 = f2_pack(0.0f, 0.0f);
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * i4);
                        const uint64_t a01 = f2_add(f2_pack(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1])), f2_pack(b4.x, b4.y));
                        const uint64_t a23 = f2_add(f2_pack(__uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3])), f2_pack(b4.z, b4.w));
                        float a0, a1, a2, a3, m0, m1, m2, m3;
                        f2_unpack(a01, a0, a1); f2_unpack(a23, a2, a3);
                        f2_unpack(f2_fma(a01, slope2, zero2), m0, m1); f2_unpack(f2_fma(a23, slope2, zero2), m2, m3);
                        f[4 * i4] = fmaxf(a0, m0); f[4 * i4 + 1] = fmaxf(a1, m1); f[4 * i4 + 2] = fmaxf(a2, m2); f[4 * i4 + 3] = fmaxf(a3, m3);
                    }
                } else if (out_nhwc) {                    // cout % 16 == 0: the chunk's 16 biases are four aligned float4
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * i4);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float a = __uint_as_float(v[4 * i4 + i]) + bb[i];
                            f[4 * i4 + i] = fmaxf(a, fmaf(a, slope, 0.0f));
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float a = __uint_as_float(v[i]) + s_bias[min(c0 + i, e_cout - 1)];
                        f[i] = fmaxf(a, fmaf(a, slope, 0.0f));
                    }
                }
                if (kF32 && out_nhwc) {
                    // fp32-storage variant: residual, output and fused max-pool in fp32 NHWC (the accumulators are never rounded)
                    if (has_resid && valid) {
                        const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e_resid) + opix * e_cout_stride + c0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const float4 r = rp[i]; f[4 * i] += r.x; f[4 * i + 1] += r.y; f[4 * i + 2] += r.z; f[4 * i + 3] += r.w; }
                    }
                    if (e_out && valid) {
                        float* op = reinterpret_cast<float*>(e_out) + opix * e_cout_stride + c0;
                        uint32_t w0[8], w1[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { w0[i] = __float_as_uint(f[i]); w1[i] = __float_as_uint(f[8 + i]); }
                        st_global_256(op, w0);
                        st_global_256(op + 8, w1);
                    }
                    if (has_pool) {
                        float m[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float o1 = xmode ? __shfl_down_sync(0xffffffffu, f[i], 1) : __shfl_xor_sync(0xffffffffu, f[i], 1);
                            const float a = fmaxf(f[i], o1);
                            m[i] = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, kTW));
                        }
                        if (valid && ((lane & 1) == (xmode ? 1 : 0)) && !(lane & kTW)) {
                            const size_t pp = ((size_t)img * (eH >> 1) + (y >> 1)) * (size_t)(eW >> 1) + (x >> 1);
                            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e_pool_out) + pp * e_cout_stride + c0);
#pragma unroll
                            for (int i = 0; i < 4; ++i) op[i] = make_float4(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]);
                        }
                    }
                    if (has_head) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float4 w4 = *reinterpret_cast<const float4*>(&s_head_w[(c0 + i) * 4]);
                            head[0] = fmaf(f[i], w4.x, head[0]); head[1] = fmaf(f[i], w4.y, head[1]);
                            head[2] = fmaf(f[i], w4.z, head[2]); head[3] = fmaf(f[i], w4.w, head[3]);
                        }
                    }
                } else if (out_nhwc) {
                    if (has_resid && valid) {
                        uint32_t rw[8];
                        ld_global_256(e_resid + opix * e_cout_stride + c0, rw);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            f[2 * i] += __uint_as_float(rw[i] << 16);
                            f[2 * i + 1] += __uint_as_float(rw[i] & 0xFFFF0000u);
                        }
                    }
                    if (has_mask && valid) {
                        // backward of the producing layer's activation, fused: g_pre = g * act'(out), out > 0 ? 1 : slope
                        uint32_t mw[8];
                        ld_global_256(e_mask + opix * e_cout_stride + c0, mw);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            f[2 * i] *= __uint_as_float(mw[i] << 16) > 0.f ? 1.f : e_mask_slope;
                            f[2 * i + 1] *= __uint_as_float(mw[i] & 0xFFFF0000u) > 0.f ? 1.f : e_mask_slope;
                        }
                    }
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                        pk[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    if (e_out && valid && !(PNNP_DBG_K & 1))
                        st_global_256(reinterpret_cast<__nv_bfloat16*>(e_out) + (is_convt ? opix * e_cout_stride : out_off) + c0, pk);     // one whole sector
                    if (has_pool) {
                        // fused nn.MaxPool2d(2): lanes l^1 hold the x-neighbour, l^16 the y-neighbour of the same tile
                        // (a warp owns two 16-pixel tile rows); max of bf16-rounded values == rounding of the max
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&pk[i]);
                            uint32_t o1 = xmode ? __shfl_down_sync(0xffffffffu, pk[i], 1) : __shfl_xor_sync(0xffffffffu, pk[i], 1);
                            a = __hmax2(a, *reinterpret_cast<__nv_bfloat162*>(&o1));
                            uint32_t cur = *reinterpret_cast<uint32_t*>(&a);
                            uint32_t o2 = __shfl_xor_sync(0xffffffffu, cur, kTW);
                            a = __hmax2(a, *reinterpret_cast<__nv_bfloat162*>(&o2));
                            pk[i] = *reinterpret_cast<uint32_t*>(&a);
                        }
                        if (valid && ((lane & 1) == (xmode ? 1 : 0)) && !(lane & kTW) && !(PNNP_DBG_K & 1)) {
                            st_global_256(reinterpret_cast<__nv_bfloat16*>(e_pool_out) + pool_off + c0, pk);
                        }
                    }
                    if (has_head) {
                        // fused 1x1 head (conv10_1, Unet.py:93): 4 dot products over this pixel's channels, fp32
                        if constexpr (kX2) {
                            uint64_t h01 = f2_pack(head[0], head[1]), h23 = f2_pack(head[2], head[3]);
                            // one base address for the chunk's 16 weight rows: indexing s_head_w[(c0 + i) * 4] per channel made the
                            // scalar form spend five 64-bit address instructions on every load (80 of its 318 per chunk)
                            const float4* hw4 = reinterpret_cast<const float4*>(s_head_w) + c0;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float4 w4 = hw4[i];
                                const uint64_t ff = f2_pack(f[i], f[i]);
                                h01 = f2_fma(ff, f2_pack(w4.x, w4.y), h01);
                                h23 = f2_fma(ff, f2_pack(w4.z, w4.w), h23);
                            }
                            f2_unpack(h01, head[0], head[1]); f2_unpack(h23, head[2], head[3]);
                        } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float4 w4 = *reinterpret_cast<const float4*>(&s_head_w[(c0 + i) * 4]);
                            head[0] = fmaf(f[i], w4.x, head[0]); head[1] = fmaf(f[i], w4.y, head[1]);
                            head[2] = fmaf(f[i], w4.z, head[2]); head[3] = fmaf(f[i], w4.w, head[3]);
                        }
                        }
                    }
                } else if (valid) {
                    float* o = reinterpret_cast<float*>(e_out);
                    const size_t plane = (size_t)eH * eW;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {           // static indices keep f[] in registers
                        if (c0 + i < e_cout) {
                            const size_t oi = ((size_t)img * e_cout + (c0 + i)) * plane + (size_t)y * eW + x;
                            o[oi] = f[i] + (head_resid ? head_resid[oi] : 0.f);
                        }
                    }
                }
            }
            if (has_head && valid) {
                const size_t plane = (size_t)eH * eW;
                size_t oi = (size_t)img * head_cout * plane + (size_t)y * eW + x;       // output 0; one plane further per output
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    if (o < head_cout) {
                        head_out[oi] = head[o] + head_bias[o] + (head_resid ? head_resid[oi] : 0.f);
                        oi += plane;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
            acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}
#undef PNNP_MODE_K
#undef PNNP_DBG_K
#undef PNNP_SWZ_K

}  // namespace pnnp
#ifndef PNNP_HOST_EMUL
#include "layout_kernels.cuh"      // nchw_f32_to_nhwc16_bf16_kernel(s), maxpool2x2_nhwc_bf16_kernel
#endif
namespace pnnp {

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
static CUtensorMapSwizzle swz_enum(int swz) {
    return swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
// activation map: NHWC bf16 viewed as (C, W, H, N); box (kc, 16, box_h, 1)
static int make_act_map(CUtensorMap* tm, const void* ptr, int n, int h, int w, int c, int kc, int box_h, int swz,
                        int stride = 1, int esz = 2, int box_w = kTileW) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * esz, (cuuint64_t)w * c * esz, (cuuint64_t)h * w * c * esz};
    // with a traversal stride s the TMA unit loads boxDim/s elements per dimension
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(box_w * stride), (cuuint32_t)(box_h * stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(tm, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(swz), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { char b[128]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r); return fail(b); }
    return 0;
}
// weight map: [taps][rows][cin] bf16 viewed as (cin, rows, taps); box (kc, umma_n, 1)
static int make_w_map(CUtensorMap* tm, const void* ptr, int taps, int rows, int cin, int kc, int umma_n, int swz, int esz = 2) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)rows, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)cin * esz, (cuuint64_t)rows * cin * esz};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)umma_n, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(tm, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(swz), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { char b[128]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled(weight) failed: %d", (int)r); return fail(b); }
    return 0;
}

static int* g_err_dev = nullptr;     // device word the kernels report pipeline time-outs into

int conv_layer_launch(const pnnp_conv_desc& d, cudaStream_t st) {
    const int mode = d.mode, cin0 = d.cin0, cin1 = d.cin1, cout = d.cout, cout_stride = d.cout_stride, n = d.n, act = d.act,
              out_mode = d.out_mode, w_rows = d.w_rows;
    int h = d.h, w = d.w;
    const void *in0 = d.in0, *in1 = d.in1, *weight = d.weight, *resid = d.resid;
    const float *bias = d.bias, *resid_nchw = d.resid_nchw;
    void* out = d.out;
    if (!in0 || !weight || (!out && !d.head_out)) return fail("conv: null pointer");
    if (cin0 % 16 || cin1 % 16) return fail("conv: channel counts must be multiples of 16");
    const int nsrc = in1 ? 2 : 1;
    // fp32-storage variant (kind::tf32): 4-byte elements, so a 128-byte swizzle span holds 32 channels
    const bool f32 = d.io_f32 != 0;
    const int esz = f32 ? 4 : 2;
    if (f32 && (d.mask || mode == MODE_CONV2S2))
        return fail("conv: the fp32-storage (tf32) variant covers the inference layers only (no activation mask, no dgrad modes)");
    int kc = f32 ? 32 : 64;
    while (kc > 16 && ((cin0 % kc) || (nsrc > 1 && (cin1 % kc)))) kc >>= 1;
    int umma_n, n_tiles;
    if (mode == MODE_CONVT) {
        // one GEMM with N = 4*cout columns ((a*2+b)*cout + co), tiled by <= 256 columns
        if (cout % 16) return fail("convT: cout must be a multiple of 16");
        umma_n = std::min(4 * cout, 256);
        if ((4 * cout) % umma_n) return fail("convT: 4*cout must be <= 256 or a multiple of 256");
        n_tiles = 4 * cout / umma_n;
    }
    else if (mode == MODE_CONV3X) {
        // x-shift folded into N: columns (dx, co), N = 3*cout
        if (cout % 16 || 3 * cout > 256) return fail("conv3x: cout must be a multiple of 16 and <= 80");
        umma_n = 3 * cout; n_tiles = 1;
    }
    else {
        const int cpad = (cout + 15) / 16 * 16;
        umma_n = std::min(cpad, 256);
        if (cpad % umma_n) return fail("conv: cout must be <= 256 or a multiple of 256");
        n_tiles = cpad / umma_n;
    }
    if (d.pool_out && (mode == MODE_CONVT || out_mode != OUT_NHWC_BF16 || (h & 1) || (w & 1)))
        return fail("conv: fused max-pool needs an NHWC bf16 output with even h, w");
    if (d.head_out && (n_tiles != 1 || out_mode != OUT_NHWC_BF16 || d.head_cout < 1 || d.head_cout > 4 || mode == MODE_CONVT ||
                       !d.head_w || !d.head_b || cout > 64))
        return fail("conv: fused 1x1 head needs a single N tile, cout <= 64 and 1..4 head channels");
    if (out_mode == OUT_NHWC_BF16 && (cout % 16)) return fail("conv: NHWC output needs cout % 16 == 0");
    if (mode == MODE_CONVT ? (w_rows != cout) : (w_rows < n_tiles * umma_n)) return fail("conv: weight tensor has the wrong number of rows");
    const bool mode_b = mode == MODE_CONV3B;
    if (mode_b && (f32 || cout > 64 || cout % 16 || out_mode != OUT_NHWC_BF16))
        return fail("conv3b: bf16 sources, NHWC bf16 output, cout a multiple of 16 up to 64");
    const int taps = (mode == MODE_CONV3 || mode == MODE_CONV3S2 || mode_b) ? 9 : ((mode == MODE_CONVT || mode == MODE_CONV2S2) ? 4 : (mode == MODE_CONV3X ? 3 : 1));
    const bool s2 = mode == MODE_CONV3S2 || mode == MODE_CONV2S2;
    if (s2 && (nsrc > 1 || (h & 1) || (w & 1))) return fail("stride-2 conv: single source, even h and w");
    const int in_h = h, in_w = w;
    if (s2) { h /= 2; w /= 2; }        // tile over the OUTPUT grid
    const int tps = mode_b ? 9 : ((mode == MODE_CONV3 || mode == MODE_CONV3X) ? 3 : 1);
    // Super-tile variant (two M = 128 tiles per pipeline stage, see the kernel's SUP parameter).  Measured in r02 (tools/r02_sweep.sh): on by default for the MODE_CONV3 layers only; originally written as opt-in until measured
    // on a B200: PNNP_CONV_SUPER=1 -> one CTA per SM, four accumulators (two super-tiles in flight); =2 -> keeps two CTAs per SM
    // for the small-K resident-weight layers (one super-tile in flight per CTA).  Only the compile-time specialised NHWC 3x3
    // layers with N <= 128 take it (decided below, once the epilogue specialisation is known).
    // ConvTranspose2d fast path (specialised pixel-shuffle epilogue, resident weights): 31 -> 25, 39 -> 35, 62 -> 50 us on the Sony
    // frame's 512 / 256 / 128-channel layers, but 89 -> 95-99 us on the 64-channel full-resolution one, which keeps the generic path
    const bool convt_fast = variant_on("PNNP_CONVT_FAST") && (cin0 > 64 || (getenv("PNNP_CONVT_FAST") && atoi(getenv("PNNP_CONVT_FAST")) > 1));
    // super-tile: measured faster on the MODE_CONV3 layers it applies to (64->64 @712x1064: 86 -> 70 us, 32->64: 64 -> 58) and
    // slower on the x-shift-in-N layers (16->32: 99 -> 109, 64->32: 134 -> 146), which therefore stay on single tiles
    const int super_env = getenv("PNNP_CONV_SUPER") ? atoi(getenv("PNNP_CONV_SUPER")) : (mode == MODE_CONV3 ? 1 : 0);
    static const bool no_spec = getenv("PNNP_CONV_NOSPEC") != nullptr;
    const int dbg_env = getenv("PNNP_CONV_DBG") ? atoi(getenv("PNNP_CONV_DBG")) : 0;
    int epi = EPI_GENERIC;
    if (!f32 && !no_spec && (mode == MODE_CONV3 || mode == MODE_CONV3X || mode_b) && out_mode == OUT_NHWC_BF16 && !dbg_env &&
        !(d.pool_out && d.head_out) && !(d.mask && (d.pool_out || d.head_out))) {
        epi = (mode == MODE_CONV3X ? EPI_X : 0) | (d.pool_out ? EPI_POOL : 0) | (d.head_out ? EPI_HEAD : 0) | (d.mask ? EPI_MASK : 0);
        // residual: built alone (32), with the x-shift-in-N mode (33) and with x-mode + fused head (37: the last residual block of the
        // ResUnet + conv10); every other combination runs the generic epilogue
        if (d.resid) epi = (d.pool_out || d.mask || (d.head_out && mode != MODE_CONV3X && !mode_b)) ? EPI_GENERIC : (epi | EPI_RESID);
    }
    // MODE_CONV3B exists with the specialised epilogues 0 / pool / head / residual / head + residual only (the callers fall back to the
    // x-shift-in-N mode for everything else)
    if (mode_b && !(epi == 0 || epi == EPI_POOL || epi == EPI_HEAD || epi == EPI_RESID || epi == (EPI_HEAD | EPI_RESID)))
        return fail("conv3b: no kernel for this epilogue (generic / masked epilogues take MODE_CONV3 or MODE_CONV3X)");
    if (!f32 && convt_fast && mode == MODE_CONVT && out_mode == OUT_NHWC_BF16 && !d.resid && !d.mask && !d.pool_out && !d.head_out &&
        act == ACT_NONE && !no_spec && !dbg_env)
        epi = EPI_CONVT;
    // K chunk the shared-memory plan below ends up with for a given A-box height.  The super-tile variant is taken only where its
    // 18-row boxes leave the plan's K chunk (hence the order in which a pixel's products are accumulated) as the default kernel
    // has it: found on the CPU model (tests/test_device_tc_on_cpu.py) — for Cout = 128 the taller boxes evict the resident weights,
    // the chunk drops from 64 to 32 channels and outputs differ from the default kernel's in the last bf16 bit.
    auto plan_kc = [&](int box_rows) {
        const int budget = 227 * 1024 - 4096 - cout * 20, cin_all = cin0 + (nsrc > 1 ? cin1 : 0);
        int k = kc;
        const int bts = (umma_n * k * esz + 1023) / 1024 * 1024;
        const int res_bytes = (cin_all / k) * (mode == MODE_CONVT ? 1 : taps) * bts, a_only = (box_rows * kTileW * k * esz + 1023) / 1024 * 1024;
        const bool res = (mode != MODE_CONVT || convt_fast) && n_tiles == 1 && res_bytes + 3 * a_only <= budget && !getenv("PNNP_NO_RESIDENT_W");
        for (;; k >>= 1) {
            const int a_b = box_rows * kTileW * k * esz, b_ts = (umma_n * k * esz + 1023) / 1024 * 1024;
            const int st_b = ((res ? a_b : a_b + tps * b_ts) + 1023) / 1024 * 1024;
            if ((budget - (res ? res_bytes : 0)) / st_b >= 3 || k == 16 || res) return k;
        }
    };
    // (not with the residual epilogue: 64->64 + residual @712x1064 82 us on single tiles, 97 us on super-tiles; 128 channels equal)
    const bool sup_wanted = !f32 && super_env > 0 && mode != MODE_CONVT && !mode_b && epi != EPI_GENERIC && umma_n <= 128 && h > kTileH && !d.resid;
    const bool sup = sup_wanted && plan_kc(2 * kTileH + 2) == plan_kc(kTileH + 2);
    const int tile_rows = mode_b ? kTileHB : (sup ? 2 * kTileH : kTileH);
    const int box_h = (mode == MODE_CONV3 || mode == MODE_CONV3X || mode_b) ? tile_rows + 2 : kTileH;
    const int box_w = mode_b ? kBoxWB : kTileW;
    // shrink the K chunk until at least 3 pipeline stages fit
    int swz, a_bytes, b_tap_stride, stage_bytes, stages, b_resident = 0, b_res_bytes = 0;
    const int smem_budget = 227 * 1024 - 4096 - cout * 20;
    const int cin_total = cin0 + (nsrc > 1 ? cin1 : 0);
    {
        // weights resident in smem when all taps x chunks fit beside >= 3 A-only stages (single N tile, not convT)
        const int swz_r = kc * esz, bts = mode_b ? 64 * swz_r : (umma_n * swz_r + 1023) / 1024 * 1024;   // MODE_CONV3B: fixed tap stride, see issue_stage_mmas
        const int res_bytes = (cin_total / kc) * (mode == MODE_CONVT ? 1 : taps) * bts;   // convT: the four taps are N columns of one block
        const int a_only = (box_h * box_w * swz_r + 1023) / 1024 * 1024;
        // ConvTranspose2d layers with a single N tile (4 * cout <= 256) can keep their weights resident too: the ConvTranspose fast path
        // (PNNP_CONVT_FAST; default for > 64 input channels since r02 — the K = 64 layer measured slower with it)
        if ((mode != MODE_CONVT || convt_fast) && n_tiles == 1 && res_bytes + 3 * a_only <= smem_budget && !getenv("PNNP_NO_RESIDENT_W")) {
            b_resident = 1; b_res_bytes = res_bytes;
        }
    }
    for (;; kc >>= 1) {
        swz = kc * esz;
        a_bytes = box_h * box_w * swz;
        b_tap_stride = mode_b ? 64 * swz : (umma_n * swz + 1023) / 1024 * 1024;
        stage_bytes = ((b_resident ? a_bytes : a_bytes + tps * b_tap_stride) + 1023) / 1024 * 1024;
        stages = std::min(kMaxStages, (smem_budget - b_res_bytes) / stage_bytes);
        if (stages >= 3 || swz == 32 || b_resident) break;
    }
    if (stages < 2) return fail("conv: tile does not fit in shared memory");
    ConvParams p{};
    p.mode = mode; p.n_img = n; p.H = h; p.W = w;
    p.tiles_x = mode == MODE_CONV3X ? (w + kTileWX - 1) / kTileWX : (mode_b ? (w + kTileWB - 1) / kTileWB : (w + kTileW - 1) / kTileW);
    p.tiles_y = (h + tile_rows - 1) / tile_rows; p.n_tiles = n_tiles;
    p.umma_n = umma_n; p.nsrc = nsrc; p.cin0 = cin0; p.cin1 = nsrc > 1 ? cin1 : 0; p.kc = kc; p.swz = swz;
    p.cout = cout; p.cout_stride = cout_stride; p.act = act; p.out_mode = out_mode;
    p.stages = stages; p.stage_bytes = stage_bytes; p.a_bytes = a_bytes; p.b_tap_stride = b_tap_stride;
    p.b_resident = b_resident; p.b_res_bytes = b_res_bytes;
    // Small-K layers with resident weights (one or two pipeline stages per tile) are bound by the per-tile hand-shake latency of
    // the single producer / MMA warp pair (~650 cycles per tile with every unit of work switched off, r01 experiments): run TWO
    // CTAs per SM for them — half the shared memory, two accumulator buffers (<= 256 TMEM columns) and 320 threads each.
    const int ksteps_tile = (cin_total / kc) * (mode == MODE_CONV3 ? 3 : (mode == MODE_CONV3S2 ? 9 : (mode == MODE_CONV2S2 ? 4 : 1)));
    bool two_ctas = b_resident && umma_n <= 128 && ksteps_tile <= 2 && !getenv("PNNP_CONV_1CTA") && !(sup && super_env == 1);
    if (two_ctas) {
        const int half_budget = 110 * 1024 - 2048 - cout * 20;
        const int st2 = std::min(stages, (half_budget - b_res_bytes) / stage_bytes);
        if (st2 >= 3) stages = st2;
        else two_ctas = false;                               // taller (super-tile) or wider (fp32) stages may not fit twice: one CTA per SM then
    }
    const int groups = two_ctas ? 2 : ((umma_n <= 128 && !getenv("PNNP_CONV_2GROUPS")) ? 4 : 2);
    p.stages = stages;
    int tc = 32; while (tc < groups * umma_n) tc <<= 1;
    p.tmem_cols = tc; p.groups = groups;
    p.bias = bias; p.out = out; p.resid = static_cast<const __nv_bfloat16*>(resid); p.resid_nchw = resid_nchw;
    p.mask = static_cast<const __nv_bfloat16*>(d.mask); p.mask_slope = d.mask_slope;
    if (d.mask && out_mode != OUT_NHWC_BF16) return fail("conv: the activation mask needs NHWC bf16 output");
    p.pool_out = static_cast<__nv_bfloat16*>(d.pool_out);
    p.head_w = d.head_w; p.head_b = d.head_b; p.head_out = d.head_out; p.head_cout = d.head_out ? d.head_cout : 0;
    p.dbg = dbg_env;
    if (!g_err_dev) { PNNP_CUDA(cudaMalloc(&g_err_dev, sizeof(int))); PNNP_CUDA(cudaMemset(g_err_dev, 0, sizeof(int))); }
    p.err = g_err_dev;
    CUtensorMap tmA0, tmA1, tmB;
    if (int e = make_act_map(&tmA0, in0, n, in_h, in_w, cin0, kc, box_h, swz, s2 ? 2 : 1, esz, box_w)) return e;
    if (nsrc > 1) { if (int e = make_act_map(&tmA1, in1, n, h, w, cin1, kc, box_h, swz, 1, esz, box_w)) return e; }
    else tmA1 = tmA0;
    if (mode == MODE_CONVT) { if (int e = make_w_map(&tmB, weight, 1, 4 * cout, cin0, kc, umma_n, swz, esz)) return e; }
    else if (int e = make_w_map(&tmB, weight, taps, w_rows, cin0 + (nsrc > 1 ? cin1 : 0), kc, umma_n, swz, esz)) return e;
    int dev = 0, sms = 0;
    PNNP_CUDA(cudaGetDevice(&dev));
    PNNP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int total_tiles = n * p.tiles_y * p.tiles_x * n_tiles;
    const int grid = std::min(total_tiles, two_ctas ? 2 * sms : sms);
    const size_t smem = (size_t)stages * stage_bytes + b_res_bytes + 1024 /*align slack*/ + (2 * kMaxStages + 2 * kMaxGroups + 2) * 8 + 16 + (size_t)cout * 4 * 5 + 64;
    if (smem > 227 * 1024) return fail("conv: shared memory budget exceeded");
    static bool attr_done = false;
    // (taps per stage, K16 slices, epilogue specialisation); specialised epilogues exist for the 3-taps-per-stage shapes
#define PNNP_SPEC_EPI(X, T, K) X(T, K, 0) X(T, K, 1) X(T, K, 2) X(T, K, 3) X(T, K, 4) X(T, K, 5) X(T, K, 32) X(T, K, 33) X(T, K, 37)
#define PNNP_FOR_EACH_CONV_VARIANT(X) X(3, 1, -1) X(3, 2, -1) X(3, 4, -1) X(1, 1, -1) X(1, 2, -1) X(1, 4, -1) \
    PNNP_SPEC_EPI(X, 3, 1) PNNP_SPEC_EPI(X, 3, 2) PNNP_SPEC_EPI(X, 3, 4) X(3, 1, 8) X(3, 2, 8) X(3, 4, 8) X(3, 1, 9) X(3, 2, 9) X(3, 4, 9) \
    X(1, 1, 16) X(1, 2, 16) X(1, 4, 16)
#define PNNP_FOR_EACH_SUPER_VARIANT(X) PNNP_SPEC_EPI(X, 3, 1) PNNP_SPEC_EPI(X, 3, 2) PNNP_SPEC_EPI(X, 3, 4) X(3, 1, 8) X(3, 2, 8) X(3, 4, 8) \
    X(3, 1, 9) X(3, 2, 9) X(3, 4, 9)        /* 9 = x-mode + activation mask: the data gradients of the 32-channel layers (r02) */
    if (!attr_done) {
#define X(T, K, E) PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        PNNP_FOR_EACH_CONV_VARIANT(X)
#undef X
        attr_done = true;
    }
    // the variant instantiations are configured on first use (with every switch forced to 0 a run loads and configures exactly
    // the kernels it did when it was measured
    const bool pdl = variant_on("PNNP_CONV_PDL");
    // packed-pair epilogue arithmetic: built for the specialised 3x3 epilogues, alone (VAR 4), with PDL (VAR 6) or with both other
    // switches (VAR 7)
    const bool x2 = variant_on("PNNP_CONV_F32X2") && epi != EPI_GENERIC && epi != EPI_CONVT &&
                    (mode == MODE_CONV3 || mode == MODE_CONV3X) && (pdl || !sup);
    static bool attr_optin_done = false;
    if ((sup || pdl || x2) && !attr_optin_done) {
#define X(T, K, E) PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                   PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        PNNP_FOR_EACH_SUPER_VARIANT(X)
#undef X
#define X(T, K, E) PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                   PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                   PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        PNNP_FOR_EACH_SUPER_VARIANT(X)
#undef X
#define X(T, K, E) PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        PNNP_FOR_EACH_CONV_VARIANT(X)
#undef X
        attr_optin_done = true;
    }
    const int k16s = swz / 32;                    // 32-byte MMA slices per K chunk (16 bf16 or 8 tf32 elements each)
    // Programmatic dependent launch (default since r02): the kernel waits (griddepcontrol.wait) before its first global access
    cudaLaunchAttribute pdl_attr[1];
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t pdl_cfg = {};
    pdl_cfg.gridDim = dim3((unsigned)grid); pdl_cfg.blockDim = dim3((unsigned)(64 + 128 * groups)); pdl_cfg.dynamicSmemBytes = smem;
    pdl_cfg.stream = st; pdl_cfg.attrs = pdl_attr; pdl_cfg.numAttrs = 1;
    bool launched = false;
    if (f32) {
#ifdef PNNP_HOST_EMUL
        return fail("conv: the fp32-storage (tf32) variant is not modelled on the host");
#else
        static bool attr_f32_done = false;
#define PNNP_FOR_EACH_F32_VARIANT(X) X(3, 1, -1) X(3, 2, -1) X(3, 4, -1) X(1, 1, -1) X(1, 2, -1) X(1, 4, -1)
        if (!attr_f32_done) {
#define X(T, K, E) PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            PNNP_FOR_EACH_F32_VARIANT(X)
#undef X
            attr_f32_done = true;
        }
#define X(T, K, E) if (!launched && tps == T && k16s == K) { PNNP_CONV_KLAUNCH(T, K, E, 8); launched = true; }
        PNNP_FOR_EACH_F32_VARIANT(X)
#undef X
        if (!launched) return fail("conv: no fp32-storage kernel variant for this (taps per stage, K chunk)");
        count_launch();
        PNNP_CUDA(cudaGetLastError());
        return 0;
#endif
    }
    if (mode_b) {
        // MODE_CONV3B instantiations: (9 taps per stage, K16 slices 1 / 2 / 4, epilogue 0 / pool / head / residual / head + residual), plain,
        // with PDL, with packed-pair arithmetic, with both (the default)
#define PNNP_FOR_EACH_B_VARIANT(X) X(9, 1, 0) X(9, 1, 2) X(9, 1, 4) X(9, 1, 32) X(9, 1, 36) X(9, 2, 0) X(9, 2, 2) X(9, 2, 4) X(9, 2, 32) X(9, 2, 36) \
                                   X(9, 4, 0) X(9, 4, 2) X(9, 4, 4) X(9, 4, 32) X(9, 4, 36)
        static bool attr_b_done = false;
        if (!attr_b_done) {
#define X(T, K, E) PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                   PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                   PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                   PNNP_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<T, K, E, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            PNNP_FOR_EACH_B_VARIANT(X)
#undef X
            attr_b_done = true;
        }
        const bool x2b = variant_on("PNNP_CONV_F32X2");
#define X(T, K, E) if (!launched && k16s == K && epi == E) { \
        if (x2b && pdl) PNNP_CUDA(cudaLaunchKernelEx(&pdl_cfg, conv_gemm_tc_kernel<T, K, E, 6>, tmA0, tmA1, tmB, p)); \
        else if (x2b) PNNP_CONV_KLAUNCH(T, K, E, 4); \
        else if (pdl) PNNP_CUDA(cudaLaunchKernelEx(&pdl_cfg, conv_gemm_tc_kernel<T, K, E, 2>, tmA0, tmA1, tmB, p)); \
        else PNNP_CONV_KLAUNCH(T, K, E, 0); \
        launched = true; }
        PNNP_FOR_EACH_B_VARIANT(X)
#undef X
        if (!launched) return fail("conv3b: no kernel variant for this (K chunk, epilogue)");
        count_launch();
        PNNP_CUDA(cudaGetLastError());
        return 0;
    }
    if (sup && (groups & 1)) return fail("conv: the super-tile variant needs an even number of accumulator buffers (internal)");
    if (x2) {
#define X(T, K, E) if (!launched && tps == T && k16s == K && epi == E) { \
        if (sup) PNNP_CUDA(cudaLaunchKernelEx(&pdl_cfg, conv_gemm_tc_kernel<T, K, E, 7>, tmA0, tmA1, tmB, p)); \
        else if (pdl) PNNP_CUDA(cudaLaunchKernelEx(&pdl_cfg, conv_gemm_tc_kernel<T, K, E, 6>, tmA0, tmA1, tmB, p)); \
        else PNNP_CONV_KLAUNCH(T, K, E, 4); \
        launched = true; }
        PNNP_FOR_EACH_SUPER_VARIANT(X)
#undef X
    }
    if (sup && !launched) {
#define X(T, K, E) if (!launched && tps == T && k16s == K && epi == E) { \
        if (pdl) PNNP_CUDA(cudaLaunchKernelEx(&pdl_cfg, conv_gemm_tc_kernel<T, K, E, 3>, tmA0, tmA1, tmB, p)); \
        else PNNP_CONV_KLAUNCH(T, K, E, 1); \
        launched = true; }
        PNNP_FOR_EACH_SUPER_VARIANT(X)
#undef X
        if (!launched) return fail("conv: no super-tile kernel variant for this (K chunk, epilogue)");
    }

#define X(T, K, E) if (!launched && tps == T && k16s == K && epi == E) { \
        if (pdl) PNNP_CUDA(cudaLaunchKernelEx(&pdl_cfg, conv_gemm_tc_kernel<T, K, E, 2>, tmA0, tmA1, tmB, p)); \
        else PNNP_CONV_KLAUNCH(T, K, E, 0); \
        launched = true; }
    PNNP_FOR_EACH_CONV_VARIANT(X)
#undef X
    if (!launched) return fail("conv: no kernel variant for this (taps per stage, K chunk)");
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pnnp

using namespace pnnp;

extern "C" int pnnp_conv2d_tc(int mode, const void* in0, int cin0, const void* in1, int cin1, const void* weight, int w_rows,
                              const float* bias, void* out, int cout, int cout_stride, int n, int h, int w, int act,
                              int out_mode, const void* resid, const float* resid_nchw, void* stream) {
    pnnp_conv_desc d{};
    d.mode = mode; d.act = act; d.out_mode = out_mode; d.n = n; d.h = h; d.w = w;
    d.in0 = in0; d.cin0 = cin0; d.in1 = in1; d.cin1 = cin1; d.weight = weight; d.w_rows = w_rows; d.bias = bias;
    d.out = out; d.cout = cout; d.cout_stride = cout_stride; d.resid = resid; d.resid_nchw = resid_nchw;
    return conv_layer_launch(d, (cudaStream_t)stream);
}

extern "C" int pnnp_conv2d_tc_ex(const pnnp_conv_desc* desc, void* stream) {
    if (!desc) return fail("conv2d_tc_ex: null descriptor");
    return conv_layer_launch(*desc, (cudaStream_t)stream);
}

extern "C" int pnnp_conv_pipeline_error(void) {
    int v = 0;
    if (g_err_dev) { cudaMemcpy(&v, g_err_dev, sizeof(int), cudaMemcpyDeviceToHost); if (v) cudaMemset(g_err_dev, 0, sizeof(int)); }
    return v;
}

#ifndef PNNP_HOST_EMUL
extern "C" int pnnp_nchw_to_nhwc16(const float* in, void* out, int n, int c, int h, int w, float scale, void* stream) {
    if (!in || !out || c > 16) return fail("nchw_to_nhwc16: bad arguments");
    const size_t total = (size_t)n * h * w;
    if (variant_on("PNNP_IN_V2") && (((size_t)h * w) & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
        const int blocks4 = (int)std::min<size_t>((total / 4 + 255) / 256, 148 * 8);
        nchw_f32_to_nhwc16_bf16_x4_kernel<<<blocks4, 256, 0, (cudaStream_t)stream>>>(in, static_cast<__nv_bfloat16*>(out), n, c, h, w, scale);
        count_launch();
        PNNP_CUDA(cudaGetLastError());
        return 0;
    }
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    nchw_f32_to_nhwc16_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, static_cast<__nv_bfloat16*>(out), n, c, h, w, scale);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_nchw_to_nhwc16_f32(const float* in, float* out, int n, int c, int h, int w, void* stream) {
    if (!in || !out || c > 16) return fail("nchw_to_nhwc16_f32: bad arguments");
    const size_t total = (size_t)n * h * w;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    nchw_f32_to_nhwc16_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n, c, h, w);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pnnp_maxpool2x2_nhwc(const void* in, void* out, int n, int h, int w, int c, void* stream) {
    if (!in || !out || (c % 8) || (h & 1) || (w & 1)) return fail("maxpool2x2: needs even h, w and c % 8 == 0");
    const size_t total = (size_t)n * (h / 2) * (w / 2) * (c / 8);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    maxpool2x2_nhwc_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16*>(in),
                                                                         static_cast<__nv_bfloat16*>(out), n, h, w, c);
    count_launch();
    PNNP_CUDA(cudaGetLastError());
    return 0;
}
#endif  // PNNP_HOST_EMUL
