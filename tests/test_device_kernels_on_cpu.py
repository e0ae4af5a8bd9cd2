"""The grid-stride CUDA kernels of pack.cu and crop_aug.cu — the very kernel source of the GPU build (csrc/pack_kernels.cuh,
csrc/crop_kernels.cuh) — compiled for the host through tests/emul/cuda_host_shim.h and executed thread by thread on the CPU.
Kernels without shared memory or warp collectives run exactly that way, so index arithmetic, vector / scalar paths, edge handling
and per-sample arithmetic are all checked without a GPU, against the goldens of the unmodified reference and the oracle:

  P1 raw2bayer (uint16 / float32 input, vector and scalar paths), P2 bayer2raw, the dark-shading pack, D2 crop + 8-mode
  augmentation, eval_crop / eval_merge, and the white-balance gains — for several (grid, block) shapes (results must not depend
  on the launch shape).  Test infrastructure only: the product has no CPU path."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_np as O
from conftest import ROOT

EMUL = os.path.join(ROOT, "tests", "emul")
SHAPES = [(1, 1), (3, 32), (7, 5)]                     # (grid, block): one serial thread, and two shapes that interleave work
_f32p, _f64p, _u16p, _i32p = (C.POINTER(t) for t in (C.c_float, C.c_double, C.c_uint16, C.c_int))


@pytest.fixture(scope="module")
def K():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(EMUL, "_build", "libkernels_host.so")
    srcs = [os.path.join(EMUL, "kernels_host.cpp"), os.path.join(EMUL, "cuda_host_shim.h")] + \
           [os.path.join(ROOT, "pnnp_b200", "csrc", f) for f in ("pack_kernels.cuh", "pack_core.cuh", "crop_kernels.cuh", "layout_kernels.cuh", "ssim_core.cuh", "copy_kernels.cuh", "actbwd_core.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-strict-aliasing", "-shared", "-fPIC", "-o", out, srcs[0]],
                       check=True)
    return C.CDLL(out)


def _p(a, t):
    return a.ctypes.data_as(t)


def _aligned(shape, dtype):
    """16-byte aligned array (the vector paths use 128-bit accesses)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    buf = np.empty(n + 16, np.uint8)
    off = (-buf.ctypes.data) % 16
    return buf[off:off + n].view(dtype).reshape(shape)


def _pack(K, raw, wp, bl, norm=True, clip=False, bias=(0, 0, 0, 0), vec=None, shape=(1, 1)):
    H, W = raw.shape[-2:]
    n = raw.shape[0] if raw.ndim == 3 else 1
    src = _aligned(raw.shape, raw.dtype)
    src[...] = raw
    out = _aligned((n, 4, H // 2, W // 2), np.float32)
    black = (C.c_double * 4)(*[float(b) + float(bl) for b in bias])
    if vec is None:
        vec = (W // 2) % 4 == 0 and (W * raw.dtype.itemsize) % 16 == 0
    fn, t = (K.emul_pack_norm_u16, _u16p) if raw.dtype == np.uint16 else (K.emul_pack_norm_f32, _f32p)
    assert fn(_p(src, t), _p(out, _f32p), n, H, W, C.c_double(wp), black, int(norm), int(clip), int(vec), *shape) == 0
    return out if raw.ndim == 3 else out[0]


def test_pack_kernels_vs_reference_goldens_and_oracle(K, golden):
    g = golden("pack")
    raw = g["rand_raw"]
    for shape in SHAPES:
        for vec in ((False, True) if (raw.shape[1] // 2) % 4 == 0 else (False,)):
            assert _pack(K, raw, 16383, 512, vec=vec, shape=shape).tobytes() == g["rand_packed"].tobytes()
            assert _pack(K, raw, 16383, 512, clip=True, bias=(1, -2, 3, 0), vec=vec, shape=shape).tobytes() == g["rand_packed_bias"].tobytes()
            assert _pack(K, raw, 16383, 512, norm=False, vec=vec, shape=shape).tobytes() == g["rand_packed_nonorm"].tobytes()
    rs = np.random.RandomState(1)
    for (H, W) in ((2, 2), (6, 10), (8, 16), (14, 40), (32, 64)):                  # ragged and vector-friendly, batched
        r16 = rs.randint(0, 16384, size=(3, H, W)).astype(np.uint16)
        want = np.stack([O.raw2bayer(f, 16383, 512, True, True) for f in r16])
        for shape in SHAPES:
            assert _pack(K, r16, 16383, 512, clip=True, shape=shape).tobytes() == want.tobytes()
            assert _pack(K, r16, 16383, 512, clip=True, vec=False, shape=shape).tobytes() == want.tobytes()
        rf = (r16[0].astype(np.float32) + rs.rand(H, W).astype(np.float32))          # float input (dark-corrected frames)
        assert _pack(K, rf, 1023, 64, shape=(3, 32)).tobytes() == O.raw2bayer(rf, 1023, 64, True, False).tobytes()


def test_pack_reciprocal_form_equals_the_division_on_every_code(K):
    """pack_core.cuh norm_one_d<RCP>: q = x r, q + (x - span q) r against the reference's float64 division (utils/isp_ops.py:92-96)
    on EVERY sensor code of both cameras through the vector kernel (the form pack.cu launches), with per-plane bias offsets, on
    float32 samples including infinities / NaN, and the fallback to the division for a divisor whose significand is all ones."""
    for wp, bl, ncodes in ((16383, 512, 16384), (1023, 64, 1024), (65535, 0, 65536), (4095, 240, 4096)):
        codes = np.arange(ncodes, dtype=np.uint16)
        raw = np.concatenate([codes, codes[::-1]]).reshape(2, -1)                   # every code on an even and on an odd row
        raw = np.concatenate([raw, np.roll(raw, 1, axis=1)], axis=0)                # ... and in even and odd columns
        assert (raw.shape[1] // 2) % 4 == 0
        for bias in ((0, 0, 0, 0), (1, -2, 3, 0), (7, 7, -7, 11)):
            for clip in (False, True):
                want = O.raw2bayer(raw, wp, bl, True, clip, np.array(bias))
                assert _pack(K, raw, wp, bl, clip=clip, bias=bias, vec=True, shape=(3, 64)).tobytes() == want.tobytes(), (wp, bl, bias, clip)
    rs = np.random.RandomState(7)
    rf = (rs.rand(8, 64).astype(np.float32) * 70000 - 2000).astype(np.float32)
    rf[0, :6] = [np.inf, -np.inf, np.nan, 0.0, -0.0, 3.4e38]
    with np.errstate(invalid="ignore"):
        want = O.raw2bayer(rf, 16383, 512, True, False)
    got = _pack(K, rf, 16383, 512, vec=True, shape=(2, 32))
    assert got.tobytes() == want.tobytes() or (np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]))
    wp_ones = float(np.nextafter(np.float64(2.0), 0)) * 8192            # wp - 0 = 1.11...1b x 2^13: the launcher falls back to the division
    fin = np.ascontiguousarray(rf[2:4])
    assert _pack(K, fin, wp_ones, 0, vec=True, shape=(2, 32)).tobytes() == O.raw2bayer(fin, wp_ones, 0, True, False).tobytes()


def test_unpack_kernel_vs_reference_golden_and_roundtrip(K, golden):
    g = golden("pack")
    u = np.ascontiguousarray(g["unpack_in"], np.float32).reshape(-1, 4, *g["unpack_in"].shape[-2:])
    n, _, h, w = u.shape
    for shape in SHAPES:
        for vec in ((False, True) if w % 4 == 0 else (False,)):
            src = _aligned(u.shape, np.float32)
            src[...] = u
            out = _aligned((n, 2 * h, 2 * w), np.uint16)
            assert K.emul_unpack_quant(_p(src, _f32p), _p(out, _u16p), n, h, w, C.c_float(16383), C.c_float(512), int(vec), *shape) == 0
            assert np.array_equal(out.reshape(g["unpack_out"].shape), g["unpack_out"])
    rs = np.random.RandomState(2)
    raw = rs.randint(0, 16384, size=(16, 24)).astype(np.uint16)
    packed = _pack(K, raw, 16383, 512, clip=True)
    src = _aligned((1,) + packed.shape, np.float32)
    src[0] = packed
    back = _aligned((1, 16, 24), np.uint16)
    assert K.emul_unpack_quant(_p(src, _f32p), _p(back, _u16p), 1, 8, 12, C.c_float(16383), C.c_float(512), 1, 3, 32) == 0
    assert np.array_equal(back[0], np.clip(raw, 512, 16383))                        # P2(P1(x)) = clip(x, bl, wp)


def test_dark_shading_pack_kernel_vs_reference_goldens(K, golden):
    g = golden("realdata")
    raw = g["ds_raw"]
    H, W = raw.shape
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        ds = np.ascontiguousarray(g[f"ds_{tag}_map"])
        assert ds.dtype == dt
        for mean in (0, 1):
            for bd in (None, 0.3712345):
                out = _aligned((1, 4, H // 2, W // 2), np.float32)
                black = (C.c_double * 4)(512.0, 512.0, 512.0, 512.0)
                m = float(ds.mean(dtype=ds.dtype)) if mean else 0.0
                for shape in SHAPES:
                    assert K.emul_pack_norm_dark_u16(_p(raw, _u16p), ds.ctypes.data_as(C.c_void_p), int(dt == np.float64), _p(out, _f32p), 1, H, W,
                                                     C.c_double(16383), black, 1, 0, C.c_double(m), mean, C.c_double(bd or 0.0),
                                                     int(bd is not None), *shape) == 0
                    assert out[0].tobytes() == g[f"ds_{tag}_m{mean}_b{int(bd is not None)}"].tobytes(), (tag, mean, bd)


def test_crop_aug_and_tiling_kernels_vs_oracle_and_goldens(K, golden):
    rs = np.random.RandomState(3)
    frame = rs.rand(4, 40, 56).astype(np.float32)
    hs, ws, modes = [0, 3, 24, 8, 1, 17, 5, 12], [0, 9, 40, 2, 30, 11, 7, 21], list(range(8))
    want = O.random_crop(frame, hs, ws, 16, modes)
    arr = lambda v: (C.c_int * len(v))(*v)
    for shape in SHAPES:
        out = _aligned((8, 4, 16, 16), np.float32)
        assert K.emul_crop_aug(_p(frame, _f32p), _p(out, _f32p), 4, 40, 56, 16, 8, arr(hs), arr(ws), arr(modes), *shape) == 0
        assert out.tobytes() == want.tobytes()
    g = golden("tiling")
    for k in range(5):                                                              # five geometries from the unmodified reference
        c, h, w, patch, base = (int(v) for v in g[f"case{k}_geom"])
        xin = np.ascontiguousarray(g[f"case{k}_x"].reshape(c, h, w), np.float32)
        tiles_want = np.ascontiguousarray(g[f"case{k}_tiles"], np.float32)
        marked = tiles_want + np.arange(tiles_want.shape[0], dtype=np.float32).reshape(-1, 1, 1, 1)   # overlap precedence observable
        for shape in SHAPES:
            tiles = _aligned(tiles_want.shape, np.float32)
            assert K.emul_eval_crop(_p(xin, _f32p), _p(tiles, _f32p), c, h, w, patch, base, *shape) == 0
            assert tiles.tobytes() == tiles_want.tobytes(), k
            merged = _aligned((c, h, w), np.float32)
            assert K.emul_eval_merge(_p(marked, _f32p), _p(merged, _f32p), c, h, w, patch, base, *shape) == 0
            assert merged.tobytes() == np.ascontiguousarray(g[f"case{k}_merged"], np.float32).tobytes(), k
            assert K.emul_eval_merge(_p(tiles, _f32p), _p(merged, _f32p), c, h, w, patch, base, *shape) == 0
            assert merged.tobytes() == xin.tobytes()                                # crop -> merge = identity


@pytest.mark.parametrize("tag", ["wb32", "wb64", "wbpy"])
def test_wb_gains_kernel_vs_reference_goldens(K, golden, tag):
    """The white-balance kernel (written after round 1's GPU budget was spent) fed what crops.wb_jitter hands to the ABI."""
    g = golden("wb_jitter")
    wb = [float(v) for v in g[f"{tag}_wb"]] if tag == "wbpy" else g[f"{tag}_wb"]
    eff = {0: np.asarray(wb[0] / g[f"{tag}_red"]), 2: np.asarray(wb[2] / g[f"{tag}_blue"])}
    kind, gain = [0] * 4, [1.0] * 4
    for ch, e in eff.items():
        kind[ch], gain[ch] = (2 if e.dtype == np.float64 else 1), float(e.reshape(-1)[0])
    for shape in SHAPES:
        x = _aligned(g["base"].shape, np.float32)
        x[...] = g["base"]
        n, c, h, w = x.shape
        assert K.emul_wb_gains(_p(x, _f32p), n, c, h, w, C.c_float(float(g[f"{tag}_rgb"][0])), (C.c_int * 4)(*kind),
                               (C.c_double * 4)(*gain), *shape) == 0
        assert x.tobytes() == g[f"{tag}_out"].tobytes()


def test_input_layout_and_pool_kernels(K):
    """Network-input conversion NCHW fp32 -> NHWC16 bf16 (default kernel == torch's bf16 rounding; the opt-in four-pixel kernel ==
    the default one bit for bit) and the 2x2 max-pool on NHWC bf16 == F.max_pool2d."""
    import torch
    import torch.nn.functional as F
    rs = np.random.RandomState(5)
    for (n, c, h, w), scale in (((2, 4, 6, 10), 1.0), ((1, 4, 16, 24), 0.37), ((1, 12, 8, 8), 1.0), ((3, 1, 2, 2), 2.5)):
        x = (rs.standard_normal((n, c, h, w)) * 3).astype(np.float32)
        xin = _aligned(x.shape, np.float32)
        xin[...] = x
        want = torch.zeros((n, h, w, 16), dtype=torch.bfloat16)
        want[..., :c] = (torch.from_numpy(x) * scale).permute(0, 2, 3, 1).to(torch.bfloat16)
        want_bits = want.view(torch.int16).numpy().view(np.uint16)
        outs = {}
        for v2 in (0, 1):
            for shape in SHAPES:
                out = _aligned((n, h, w, 16), np.uint16)
                out[...] = 0xFFFF
                assert K.emul_nchw_to_nhwc16(_p(xin, _f32p), _p(out, _u16p), n, c, h, w, C.c_float(scale), v2, *shape) == 0
                assert np.array_equal(out, want_bits), (v2, shape)
    x = torch.from_numpy(rs.standard_normal((2, 16, 12, 20)).astype(np.float32)).to(torch.bfloat16)      # NCHW
    nhwc = _aligned((2, 12, 20, 16), np.uint16)
    nhwc[...] = x.permute(0, 2, 3, 1).contiguous().view(torch.int16).numpy().view(np.uint16)
    want = F.max_pool2d(x.float(), 2).to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().view(torch.int16).numpy().view(np.uint16)
    for shape in SHAPES:
        out = _aligned((2, 6, 10, 16), np.uint16)
        assert K.emul_maxpool2x2_nhwc(_p(nhwc, _u16p), _p(out, _u16p), 2, 12, 20, 16, *shape) == 0
        assert np.array_equal(out, want)


def test_separable_ssim_phases_vs_oracle(K):
    """The opt-in separable SSIM / squared-error pass (csrc/ssim_core.cuh), phase by phase as the device runs it between its
    __syncthreads(): PSNR / SSIM from its partial sums == the oracle's restatement of skimage's defaults, including frames that
    are not multiples of the 32 x 16 tile, scaling + clamping of the estimate, and the IlluminanceCorrect gain."""
    import math
    rs = np.random.RandomState(7)
    for (n, c, h, w), scale, use_gain in (((1, 4, 40, 56), 1.0, 0), ((2, 3, 23, 37), 1.7, 0), ((1, 4, 64, 96), 1.0, 1), ((1, 1, 7, 7), 1.0, 0)):
        hr = rs.rand(n, c, h, w).astype(np.float32)
        dn = np.clip(hr + rs.standard_normal((n, c, h, w)).astype(np.float32) * 0.08, -0.2, 1.3).astype(np.float32) / np.float32(scale)
        sums = np.zeros((n, 3 + c), np.float64)
        est = []
        for f in range(n):
            p = np.clip(dn[f] * np.float32(scale), 0, 1).astype(np.float32)
            if use_gain:
                m = hr[f] != 1.0
                num, den = float((p[m].astype(np.float64) * hr[f][m]).sum()), float((p[m].astype(np.float64) ** 2).sum())
                sums[f, 0], sums[f, 1] = num, den
                p = (np.float32(num) / np.float32(den)) * p
            est.append(p)
        assert K.emul_ssim_mse_v2(_p(np.ascontiguousarray(dn), _f32p), _p(hr, _f32p), n, c, h, w, C.c_float(scale), use_gain,
                                  _p(sums, _f64p)) == 0
        for f in range(n):
            X, Y = O.tensor2im(hr[f:f + 1]), O.tensor2im(est[f][None])                         # target, estimate as HWC x255
            mse = sums[f, 2] / (c * h * w)
            assert abs(10.0 * math.log10(255.0 ** 2 / mse) - O.psnr(X, Y)) < 1e-9
            ssim = sums[f, 3:3 + c].sum() / (c * (h - 6) * (w - 6))
            assert abs(ssim - O.ssim(X, Y)) < 1e-9, (ssim, O.ssim(X, Y))


def test_strided_copy_kernels_pack_weights_like_numpy(K):
    """The batched strided copy / cast of the training step (weight packing incl. the 180-degree filter flip through negative source
    strides, bf16 casts) — default kernel and the opt-in 32-bit-index form — against NumPy's own strided views."""
    import torch
    from pnnp_b200 import _lib
    rs = np.random.RandomState(11)
    w = rs.standard_normal((24, 16, 3, 3)).astype(np.float32)              # [co][ci][ky][kx]
    # forward layout [tap][co][ci] (bf16), flipped + transposed data-gradient layout [tap'][ci][co] (bf16), plain fp32 transpose
    jobs = [((3, 3, 24, 16), w.transpose(2, 3, 0, 1), True), ((3, 3, 16, 24), w[:, :, ::-1, ::-1].transpose(2, 3, 1, 0), True),
            ((16, 24, 3, 3), w.transpose(1, 0, 2, 3), False)]
    for v2 in (0, 1):
        descs = (_lib.CopyDesc * len(jobs))()
        outs = []
        for d, (dims, view, to_bf16) in zip(descs, jobs):
            assert view.shape == dims
            out = np.zeros(dims, np.uint16 if to_bf16 else np.float32)
            outs.append(out)
            first = view[0, 0, 0, 0:1]                                    # address of logical index 0 (negative strides start late in memory)
            d.src, d.dst, d.dst_bf16 = first.ctypes.data, out.ctypes.data, int(to_bf16)
            for k in range(4):
                d.dim[k], d.sstride[k], d.dstride[k] = dims[k], view.strides[k] // 4, out.strides[k] // out.itemsize
        for shape in SHAPES:
            for o in outs:
                o[...] = 0
            assert K.emul_strided_copy_batch(descs, len(jobs), v2, *shape) == 0
            for (dims, view, to_bf16), o in zip(jobs, outs):
                if to_bf16:
                    want = torch.from_numpy(np.ascontiguousarray(view)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
                else:
                    want = np.ascontiguousarray(view)
                assert np.array_equal(o, want), (v2, shape, dims)


@pytest.mark.parametrize("act_kind", [0, 1, 2])
def test_act_backward_v2_phases_vs_torch(K, act_kind):
    """The opt-in activation-backward + bias-gradient kernel (csrc/actbwd_core.cuh), phase by phase: the in-place gradient update is
    bit-identical to torch's bf16 arithmetic (g * act'(out), rounded to bf16), the bias gradient is the per-channel sum."""
    import torch
    rs = np.random.RandomState(13 + act_kind)
    for (pixels, c, nblocks) in ((300, 16, 1), (1000, 32, 3), (257, 64, 2), (64, 512, 5)):
        g = torch.from_numpy(rs.standard_normal((pixels, c)).astype(np.float32)).to(torch.bfloat16)
        out = torch.from_numpy(rs.standard_normal((pixels, c)).astype(np.float32)).to(torch.bfloat16)
        if act_kind == 1:
            want = (g.float() * torch.where(out.float() > 0, 1.0, 0.2)).to(torch.bfloat16)
        elif act_kind == 2:
            want = torch.where(out.float() > 0, g.float(), torch.zeros(())).to(torch.bfloat16)
        else:
            want = g.clone()
        gb = _aligned((pixels, c), np.uint16)
        gb[...] = g.view(torch.int16).numpy().view(np.uint16)
        ob = _aligned((pixels, c), np.uint16)
        ob[...] = out.view(torch.int16).numpy().view(np.uint16)
        dbias = np.full(c, 0.5, np.float32)                                    # accumulates on top of what is there
        assert K.emul_act_bwd_bias_v2(_p(gb, _u16p), _p(ob, _u16p), _p(dbias, _f32p), pixels * (c // 8), c, act_kind, nblocks) == 0
        assert np.array_equal(gb, want.view(torch.int16).numpy().view(np.uint16))
        assert np.allclose(dbias - 0.5, want.float().sum(0).numpy(), rtol=1e-4, atol=1e-3)
