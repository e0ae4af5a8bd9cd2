"""WHOLE warp-collective kernels of the noise synthesis (pnnp_b200/csrc/noise_kernels.cuh) compiled for the host and run, CTA by
CTA, on a lock-step fibre emulator (tests/emul/simt_host.h: every CUDA thread is a fibre; __ballot_sync / __shfl_sync /
__syncwarp / __syncthreads park the fibre until its warp / block has arrived).  The specialised BASELINE-configs[1] kernel —
sampler sorting through the shared-memory queue, shuffled row draws, prefetch, 128-bit accesses — is checked without a GPU:

  * bit-identical output AND draws to the generic kernel (128-bit and scalar paths), for ragged widths, several grid sizes and
    crop offsets: results do not depend on who computes an element;
  * its output is the reference's arithmetic (oracle restatement of process.py:593-631, pinned on the reference goldens) applied to
    the draws it recorded;
  * shard independence: a crop synthesised alone (crop_id0 = k) equals crop k of the batch.

The three MUFU approximations are libm on the host, so the DRAWS differ from a B200's in the last bits; every identity above is
between kernels / the oracle fed the same draws, exactly what the `-m gpu` tests check on the device.  Test infrastructure only."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_np as O
from conftest import ROOT
from pnnp_b200 import _lib
from pnnp_b200.noise import noise_code_bits
from pnnp_b200.noise_params import fill_row

EMUL = os.path.join(ROOT, "tests", "emul")
_f32p, _f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
GENERIC4, GENERIC1, FAST = 0, 1, 2


@pytest.fixture(scope="module")
def S():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(EMUL, "_build", "libsimt_kernels_host.so")
    srcs = [os.path.join(EMUL, f) for f in ("simt_kernels_host.cpp", "simt_host.h", "cuda_host_shim.h")] + \
           [os.path.join(ROOT, "pnnp_b200", "csrc", f) for f in ("noise_kernels.cuh", "noise_core.cuh")] + [os.path.join(ROOT, "include", "pnnp_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-strict-aliasing", "-shared", "-fPIC", "-o", out, srcs[0]], check=True)
    return C.CDLL(out)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _synth(S, clean, params, code, kernel, blocks, chain=_lib.CHAIN_NUMPY, seed=7, offset=3, crop_id0=0, post=(-np.inf, np.inf),
           ori=False, clip=False):
    n, c, h, w = clean.shape
    rows = (_lib.NoiseParamsRow * n)()
    for r, p in zip(rows, params):
        fill_row(r, p, chain == _lib.CHAIN_TORCH)
    clean = np.ascontiguousarray(clean, np.float32)
    out = np.full_like(clean, np.nan)
    d = {"counts": np.full_like(clean, np.nan), "read": np.full_like(clean, np.nan), "row_z": np.full((n, c, h, 1), np.nan, np.float32),
         "q": np.full(clean.shape, np.nan, np.float64)}
    rc = S.emul_noise_synth(_p(clean, _f32p), _p(out, _f32p), rows, n, c, h, w, C.c_uint32(noise_code_bits(code)), chain, int(ori), int(clip),
                            C.c_float(post[0]), C.c_float(post[1]), C.c_uint64(seed), C.c_uint64(offset), C.c_uint64(crop_id0), kernel, 1,
                            _p(d["counts"], _f32p), _p(d["read"], _f32p), _p(d["row_z"], _f32p), _p(d["q"], _f64p), blocks)
    assert rc == 0
    return out, d


def _crops(n, c, h, w, seed):
    rs = np.random.RandomState(seed)
    clean = (rs.rand(n, c, h, w).astype(np.float32)) ** 2
    clean[:, :, :, : w // 8] *= 0.02                                   # a dark stretch: rates below the sampler switch in every row
    np.random.seed(seed)
    return clean, [O.sample_params("SonyA7S2") for _ in range(n)]


@pytest.mark.parametrize("w,blocks,crop_id0", [(512, 1, 0), (644, 3, 5), (1028, 2, 1)])
def test_specialised_kernel_is_bit_identical_to_the_generic_kernel(S, w, blocks, crop_id0):
    clean, params = _crops(3, 4, 3, w, 11 + w)
    fast, df = _synth(S, clean, params, "pgrq", FAST, blocks, crop_id0=crop_id0)
    assert np.isfinite(fast).all()
    for kernel, b in ((GENERIC4, blocks + 1), (GENERIC1, 2)):
        if kernel == GENERIC4 and (crop_id0 * 4 * 3 * w) % 4:
            continue
        gen, dg = _synth(S, clean, params, "pgrq", kernel, b, crop_id0=crop_id0)
        for k in df:
            assert df[k].tobytes() == dg[k].tobytes(), (k, kernel)
        assert fast.tobytes() == gen.tobytes(), kernel
    # both Poisson samplers were exercised, and the counts follow the rates
    lam = clean * np.array([(p["wp"] - p["bl"]) / p["ratio"] / p["K"] for p in params], np.float64).reshape(-1, 1, 1, 1)
    assert (lam < 10).mean() > 0.1 and (lam >= 10).mean() > 0.1
    assert abs((df["counts"] - lam).mean()) < 4 * np.sqrt(lam.mean() / lam.size) + 1e-3


def test_specialised_kernel_output_is_the_reference_arithmetic_on_its_draws(S):
    clean, params = _crops(2, 4, 2, 1152, 5)
    out, d = _synth(S, clean, params, "pgrq", FAST, 2, post=(-np.inf, 1.0))
    for i, p in enumerate(params):
        want = O.noisy_obs_tail_explicit(clean[i], p, "pgrq", {k: v[i] for k, v in d.items()})
        assert np.minimum(want, np.float32(1.0)).tobytes() == out[i].tobytes()            # fused lr.clip(-inf, 1) (syn_datasets.py:339-342)
    assert (out == 1.0).any()
    # row noise: one draw per (crop, channel, row), N(0, 1)-sized; quantisation draws on the 12-bit lattice inside (-0.5, 0.5)
    assert np.abs(d["row_z"]).max() < 6 and len(np.unique(d["row_z"])) == d["row_z"].size
    assert np.abs(d["q"]).max() < 0.5 and np.array_equal(d["q"] * 8192, np.round(d["q"] * 8192))


def test_crops_do_not_depend_on_the_batch_they_are_synthesised_in(S):
    clean, params = _crops(3, 4, 2, 516, 23)
    full, _ = _synth(S, clean, params, "pgrq", FAST, 2, crop_id0=4)
    for k in range(3):
        alone, _ = _synth(S, clean[k:k + 1], params[k:k + 1], "pgrq", FAST, 1, crop_id0=4 + k)
        assert alone[0].tobytes() == full[k].tobytes()
    other, _ = _synth(S, clean, params, "pgrq", FAST, 2, crop_id0=4, offset=4)
    assert other.tobytes() != full.tobytes()                                                # another Philox offset: other draws


@pytest.mark.parametrize("code,chain", [("pgrq", _lib.CHAIN_NUMPY), ("pg", _lib.CHAIN_NUMPY), ("prq", _lib.CHAIN_TORCH), ("gr", _lib.CHAIN_NUMPY)])
def test_generic_kernel_vector_and_scalar_paths_agree_and_follow_the_oracle_tail(S, code, chain):
    clean, params = _crops(2, 4, 2, 260, 31)
    if chain == _lib.CHAIN_TORCH:
        np.random.seed(3)
        params = [O.sample_params_max("SonyA7S2", ratio=None) for _ in range(2)]
    a, da = _synth(S, clean, params, code, GENERIC4, 2, chain=chain)
    b, db = _synth(S, clean, params, code, GENERIC1, 3, chain=chain)
    assert a.tobytes() == b.tobytes()
    if chain == _lib.CHAIN_NUMPY and "p" in code:
        for i, p in enumerate(params):
            draws = {k: v[i] for k, v in da.items()}
            assert O.noisy_obs_tail_explicit(clean[i], p, code, draws).tobytes() == a[i].tobytes()


# ----------------------------------------------------------------------------------------------------------------------------
# Training-step helper kernels (csrc/train_kernels.cuh) — warp-shuffle reductions, shared-memory accumulators, block barriers —
# against torch's CPU autograd / optimiser.  bf16 tensors travel as uint16 bit patterns.
# ----------------------------------------------------------------------------------------------------------------------------
_u16p = C.POINTER(C.c_uint16)


def _bf16_bits(t):
    import torch
    return t.to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16).copy()       # a private copy: kernels work in place


def _from_bits(a):
    import torch
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


@pytest.mark.parametrize("blocks", [1, 3])
def test_l1_loss_kernel_vs_torch_autograd(S, blocks):
    """losses/base_loss.py:92-103: F.l1_loss(pred.clamp(0, 1), hr), mean reduction, and its gradient (clamp passes 0 <= p <= 1,
    sign(0) = 0)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    pred = (torch.rand((2, 4, 9, 23), generator=g) * 1.6 - 0.3)
    hr = torch.rand((2, 4, 9, 23), generator=g)
    pred.view(-1)[:6] = torch.tensor([0.0, 1.0, 0.5, -0.0, 1.0000001, 0.25])
    hr.view(-1)[:6] = torch.tensor([0.3, 0.2, 0.5, 0.0, 0.4, 0.25])                  # exact ties: zero gradient
    p = pred.clone().requires_grad_(True)
    want = F.l1_loss(p.clamp(0, 1), hr)
    want.backward()
    pn, hn = pred.numpy().copy(), hr.numpy().copy()
    gp, loss = np.full_like(pn, np.nan), C.c_double(-1.0)
    assert S.emul_l1_loss(_p(pn, _f32p), _p(hn, _f32p), _p(gp, _f32p), C.c_size_t(pn.size), C.byref(loss), blocks) == 0
    assert abs(loss.value / pn.size - want.item()) < 1e-7
    assert np.array_equal(gp, p.grad.numpy())


@pytest.mark.parametrize("cin,co,act_kind,blocks", [(32, 4, 1, 2), (16, 4, 1, 1), (64, 3, 2, 3), (8, 1, 0, 1)])
def test_head_backward_kernel_vs_torch_autograd(S, cin, co, act_kind, blocks):
    """1x1 head (conv10_1) backward fused with the feeding layer's activation derivative: data gradient (bf16, NHWC), weight and
    bias gradients of the head, bias gradient of the feeding conv — lanes of a warp are combined by shuffles, then shared / global
    accumulators."""
    import torch
    g = torch.Generator().manual_seed(cin + co)
    n, h, w = 2, 5, 13                                                   # 130 pixels: ragged against every channel-group count
    pre = torch.randn((n, h, w, cin), generator=g)
    slope = {0: 1.0, 1: 0.2, 2: 0.0}[act_kind]
    act = torch.where(pre > 0, pre, pre * slope).to(torch.bfloat16)     # the stored activation the kernel reads
    W = torch.randn((co, cin), generator=g) * 0.3
    gpred = torch.randn((n, co, h, w), generator=g)
    a32 = act.float()
    gp_pix = gpred.permute(0, 2, 3, 1)                                   # n h w co
    g_act = gp_pix.double() @ W.double()                                 # d loss / d activation
    deriv = torch.where(a32 > 0, torch.ones(()), torch.full((), slope)).double()
    g_pre = g_act * deriv
    want_dW = torch.einsum("nhwo,nhwc->oc", gp_pix.double(), a32.double())
    want_db = gp_pix.double().sum((0, 1, 2))
    want_dbp = g_pre.sum((0, 1, 2))
    act_b = _bf16_bits(act)
    gact = np.zeros_like(act_b)
    dW, db, dbp = np.full((co, cin), 0.5, np.float32), np.full(co, 0.25, np.float32), np.full(cin, -1.0, np.float32)
    Wn, gpn = W.numpy().copy(), gpred.numpy().copy()
    assert S.emul_head_bwd(_p(gpn, _f32p), _p(act_b, _u16p), _p(Wn, _f32p), _p(gact, _u16p), _p(dW, _f32p), _p(db, _f32p), _p(dbp, _f32p),
                           n, h, w, cin, co, act_kind, blocks) == 0
    got = _from_bits(gact).float().reshape(n, h, w, cin).double()
    # bf16 data gradient: within one rounding of the float64 value (fp32 FMA chain, then round to nearest even)
    assert ((got - g_pre).abs() <= g_pre.abs() * 2.0 ** -8 + 1e-30).all()
    assert (got == g_pre.float().to(torch.bfloat16).double()).float().mean() > 0.995
    assert np.allclose(dW - 0.5, want_dW.numpy(), rtol=1e-4, atol=1e-4)          # accumulates on top of what is there
    assert np.allclose(db - 0.25, want_db.numpy(), rtol=1e-4, atol=1e-4)
    assert np.allclose(dbp + 1.0, want_dbp.numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("act_kind", [0, 1, 2])
def test_act_backward_kernels_whole_vs_torch(S, act_kind):
    """The default activation-backward + bias-gradient kernel and the opt-in second form, as whole kernels: in-place gradient
    bit-identical to torch's bf16 arithmetic and to each other, bias sums per channel."""
    import torch
    rs = np.random.RandomState(3 + act_kind)
    for pixels, c, blocks in ((301, 16, 1), (999, 32, 2), (130, 64, 1), (77, 512, 3)):
        g = torch.from_numpy(rs.standard_normal((pixels, c)).astype(np.float32)).to(torch.bfloat16)
        out = torch.from_numpy(rs.standard_normal((pixels, c)).astype(np.float32)).to(torch.bfloat16)
        if act_kind == 1:
            want = (g.float() * torch.where(out.float() > 0, 1.0, 0.2)).to(torch.bfloat16)
        elif act_kind == 2:
            want = torch.where(out.float() > 0, g.float(), torch.zeros(())).to(torch.bfloat16)
        else:
            want = g.clone()
        res = []
        for fn, args in ((S.emul_act_bwd_bias, (C.c_size_t(pixels),)), (S.emul_act_bwd_bias_v2_kernel, (C.c_uint32(pixels * (c // 8)),))):
            gb, ob, dbias = _bf16_bits(g), _bf16_bits(out), np.full(c, 0.5, np.float32)
            assert fn(_p(gb, _u16p), _p(ob, _u16p), _p(dbias, _f32p), *args, c, act_kind, blocks) == 0
            assert np.array_equal(gb, _bf16_bits(want))
            assert np.allclose(dbias - 0.5, want.float().sum(0).numpy(), rtol=1e-4, atol=1e-3)
            res.append(gb)
        assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("act_kind,skip", [(0, False), (1, True), (2, True), (1, False)])
def test_maxpool_backward_kernel_vs_torch_autograd(S, act_kind, skip):
    """2x2 max-pool backward (+ skip-connection gradient, + act'(cfull)): the pooled gradient goes to the FIRST arg-max of each window
    in scan order, as torch's CPU max_pool2d does (ties are common in bf16)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(17 + act_kind)
    n, h, w, c = 2, 6, 10, 24
    cfull = (torch.randint(-3, 4, (n, c, h, w), generator=g).float() * 0.25).to(torch.bfloat16)      # few distinct values: many ties
    gp = torch.randn((n, c, h // 2, w // 2), generator=g).to(torch.bfloat16)
    gskip = torch.randn((n, c, h, w), generator=g).to(torch.bfloat16) if skip else None
    x = cfull.float().requires_grad_(True)
    F.max_pool2d(x, 2).backward(gp.float())
    tot = x.grad + (gskip.float() if skip else 0.0)
    want = tot.to(torch.bfloat16)                                        # the sum is rounded to bf16 first ...
    if act_kind:
        sl = 0.2 if act_kind == 1 else 0.0
        want = (want.float() * torch.where(cfull.float() > 0, 1.0, sl)).to(torch.bfloat16)      # ... then multiplied by act'
    nhwc = lambda t: _bf16_bits(t.permute(0, 2, 3, 1))
    gc = np.zeros((n, h, w, c), np.uint16)
    for blocks in (1, 2):
        gc[...] = 0
        assert S.emul_maxpool_bwd(_p(nhwc(gp), _u16p), _p(nhwc(cfull), _u16p), _p(nhwc(gskip), _u16p) if skip else None, _p(gc, _u16p),
                                  n, h, w, c, act_kind, blocks) == 0
        got, ref = _from_bits(gc).float(), want.permute(0, 2, 3, 1).float()
        assert torch.equal(got + 0.0, ref + 0.0)                         # +0.0: -0 and +0 compare equal either way; kept explicit


@pytest.mark.parametrize("act_kind,skip,c", [(1, True, 32), (0, False, 64), (1, True, 256)])
def test_maxpool_backward_with_bias_sums_kernel(S, act_kind, skip, c):
    """The fused form (pnnp_maxpool_bwd_bias): the routed gradient is bit-identical to the plain kernel's, and dbias receives the
    per-channel sums of the stored bf16 values on top of what it held (the separate read-only pass it replaces)."""
    import torch
    g = torch.Generator().manual_seed(5 + c)
    n, h, w = 2, 6, 10
    cfull = (torch.randint(-3, 4, (n, c, h, w), generator=g).float() * 0.25).to(torch.bfloat16)
    gp = torch.randn((n, c, h // 2, w // 2), generator=g).to(torch.bfloat16)
    gskip = torch.randn((n, c, h, w), generator=g).to(torch.bfloat16) if skip else None
    nhwc = lambda t: _bf16_bits(t.permute(0, 2, 3, 1))
    args = (_p(nhwc(gp), _u16p), _p(nhwc(cfull), _u16p), _p(nhwc(gskip), _u16p) if skip else None)
    ref = np.zeros((n, h, w, c), np.uint16)
    assert S.emul_maxpool_bwd(*args, _p(ref, _u16p), n, h, w, c, act_kind, 2) == 0
    want = _from_bits(ref).double().reshape(-1, c).sum(0).numpy()
    for blocks in (1, 2, 3):
        gc, db = np.zeros((n, h, w, c), np.uint16), np.full(c, 0.5, np.float32)
        assert S.emul_maxpool_bwd_bias(*args, _p(gc, _u16p), _p(db, _f32p), n, h, w, c, act_kind, blocks) == 0
        assert np.array_equal(gc, ref)
        assert np.allclose(db - 0.5, want, rtol=1e-5, atol=1e-4)
    assert S.emul_maxpool_bwd_bias(*args, _p(gc, _u16p), _p(db, _f32p), n, h, w, 24, act_kind, 1) == 1      # 3 channel groups do not divide 256


@pytest.mark.parametrize("on_device_state", [0, 1])
def test_adam_kernels_vs_torch_optim(S, on_device_state):
    """torch.optim.Adam defaults (betas 0.9 / 0.999, eps 1e-8, no weight decay), three steps; `gscale` = the 1 / world factor of the
    DDP gradient mean folded into the update."""
    import torch
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) * s for s in (1.0, 0.1, 3.0)]
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    p, m, v = p0.numpy().copy(), np.zeros(1000, np.float32), np.zeros(1000, np.float32)
    for step, gr in enumerate(grads, 1):
        ref.grad = gr.clone() * 0.5
        opt.step()
        gn = gr.numpy().copy()
        assert S.emul_adam(_p(p, _f32p), _p(gn, _f32p), _p(m, _f32p), _p(v, _f32p), C.c_size_t(1000), C.c_float(1e-3), C.c_float(0.9),
                           C.c_float(0.999), C.c_float(1e-8), step, C.c_float(0.5), on_device_state, 2) == 0
        assert np.allclose(p, ref.detach().numpy(), rtol=2e-6, atol=2e-7)


# ----------------------------------------------------------------------------------------------------------------------------
# Eval epilogue (csrc/eval_kernels.cuh: IlluminanceCorrect dots, squared error, SSIM map sums; shuffle + shared-memory block
# reductions, 3-D grid) and the HighBitRecovery map (csrc/hbr_kernels.cuh) — the same checks the `-m gpu` files make on a B200.
# ----------------------------------------------------------------------------------------------------------------------------
def _oracle_metrics(dn, hr, scale, correct):
    import torch
    d = torch.clamp(torch.from_numpy(dn) * scale, 0, 1)
    t = torch.from_numpy(hr)
    if correct:
        d = O.illuminance_correct(d, t)
    a, b = O.tensor2im(d.numpy()), O.tensor2im(hr)
    return O.psnr(b, a), O.ssim(b, a)


@pytest.mark.parametrize("v2", [0, 1])
@pytest.mark.parametrize("shape,correct,scale", [((1, 4, 40, 72), False, 1.0), ((2, 3, 37, 67), True, 1.0), ((1, 4, 33, 34), True, 100.0)])
def test_eval_epilogue_kernels_vs_oracle(S, golden, shape, correct, scale, v2):
    from pnnp_b200.metrics import finish_metrics
    import torch
    n, c, h, w = shape
    rs = np.random.RandomState(h)
    hr = rs.rand(*shape).astype(np.float32)
    hr[:, 0, :2, :7] = 1.0                                       # saturated pixels are excluded from the gain
    dn = ((hr * 0.9 + 0.05 * rs.randn(*shape)) / scale).astype(np.float32)
    sums = np.full((n, 3 + c), np.nan, np.float64)
    assert S.emul_eval_epilogue(_p(dn, _f32p), _p(hr, _f32p), n, c, h, w, C.c_float(scale), int(correct), _p(sums, _f64p), v2, 3) == 0
    res = finish_metrics(torch.from_numpy(sums), c, h, w)
    for i in range(n):
        p, s = _oracle_metrics(dn[i:i + 1], hr[i:i + 1], scale, correct)
        assert res[i]["PSNR"] == pytest.approx(p, abs=2e-4) and res[i]["SSIM"] == pytest.approx(s, abs=1e-6), (res[i], p, s)
    if correct and scale == 1.0:                                 # the gain itself: num / den of the reference's masked dot products
        d = np.clip(dn[0], 0, 1)
        m = hr[0] != 1
        assert sums[0, 0] / sums[0, 1] == pytest.approx(float(np.dot(d[m].astype(np.float64), hr[0][m])) / float(np.dot(d[m].astype(np.float64), d[m])), rel=1e-12)


def test_illuminance_correct_kernel_vs_reference_golden(S, golden):
    """data_process/__init__.py:162-175 on the golden of the unmodified reference: gain = num / den in float32, times the clamped prediction."""
    g = golden("eval")
    pred, src = np.ascontiguousarray(np.clip(g["pred"], 0, 1), np.float32), np.ascontiguousarray(g["src"], np.float32)
    n, c, h, w = pred.shape
    sums = np.zeros((n, 3 + c), np.float64)
    assert S.emul_eval_epilogue(_p(pred, _f32p), _p(src, _f32p), n, c, h, w, C.c_float(1.0), 1, _p(sums, _f64p), 0, 2) == 0
    gain = np.float32(sums[0, 0]) / np.float32(sums[0, 1])
    np.testing.assert_allclose(gain * pred, g["corrected"], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("k,cam,code,iso", ((0, "SonyA7S2", "pgrq", 3200), (1, "IMX686", "prq", 6400)))
def test_hbr_map_kernel_vs_reference_goldens(S, golden, k, cam, code, iso):
    """HighBitRecovery.map (process.py:726-751) fed the reference's own uniforms: float64 quantile (libm here, CUDA's math library on
    the device, SciPy in the reference) rounded to float32 — equal to the reference except for isolated last-bit differences; and
    the Philox path re-draws every sample inside its quantisation cell."""
    from pnnp_b200.real_preproc import HighBitRecovery
    g = golden("realdata")
    np.random.seed(100 + k)
    hb = HighBitRecovery(camera_type=cam, noise_code=code)
    hb.get_lut([iso], blc_mean=None)
    lut = hb.lut[iso]
    p = lut["param"]
    span = float(p["wp"] - p["bl"])
    tukey = "g" in code
    tab = np.ascontiguousarray(lut["_table"])

    def run(data, norm, rand, seed=0, offset=0):
        data = np.ascontiguousarray(data, np.float32)
        out, r_out = np.full_like(data, np.nan), np.full(data.shape, np.nan, np.float64)
        assert S.emul_hbr_map(_p(data, _f32p), _p(out, _f32p), C.c_size_t(data.size), _p(tab[0], _f64p), _p(tab[1], _f64p), int(lut["low"]),
                              int(lut["high"]), int(data.max() <= 1), int(norm), C.c_float(span), C.c_float(p["bl"]), int(tukey),
                              C.c_double(p["lam"]), C.c_double(lut["bias"]), C.c_double(p["sigTL"] if tukey else p["sigGs"]),
                              _p(rand, _f64p), C.c_uint64(seed), C.c_uint64(offset), C.c_uint64(0), _p(r_out, _f64p), 3) == 0
        return out, r_out

    data, rand = g[f"hbr{k}_data"], np.ascontiguousarray(g[f"hbr{k}_rand"], np.float64)
    for d_in, norm, want in ((data, True, g[f"hbr{k}_out"]), (data * np.float32(span), False, g[f"hbr{k}_out_dn"])):
        got, _ = run(d_in, norm, rand)
        assert got.tobytes() == want.tobytes()                         # glibc's libm here: bit-exact
    from scipy import stats
    cell = np.full((64, 64), 2.0 / span, np.float32)                     # every sample in the DN cell x = 2
    out, u = run(cell, False, None, seed=11, offset=5)
    assert stats.kstest(u.ravel(), "uniform").pvalue > 1e-3
    x = out.ravel().astype(np.float64) - p["bl"]
    assert x.min() >= 1.5 - 1e-3 and x.max() <= 2.5 + 1e-3
    assert np.array_equal(run(cell, False, np.ascontiguousarray(u))[0], out)    # same arithmetic with the draws replayed
