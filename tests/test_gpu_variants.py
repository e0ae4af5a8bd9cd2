"""Kernel variants written at the end of round 1 and measured on a B200 at the start of round 2 (tools/r02_sweep.sh ->
profiles/r02_sweep_summary.txt).  The ones that were bit-equal AND faster are now the defaults (NAME=0 switches one off): PDL,
packed-pair epilogues, the ConvTranspose2d fast path (> 64 input channels), the super-tile on the MODE_CONV3 layers, the 4-pixel
input conversion, the register-resident wgrad loops, the 32-bit copy / activation-backward kernels and the separable SSIM sums.
These tests keep proving the equalities the promotion rests on: every variant against the round-1 kernel it replaced (all
switches forced to 0 by the fixture below), bit for bit where the arithmetic is the same.

* PNNP_CONV_SUPER=1|2 — super-tile conv kernel (conv_tc.cu, template parameter SUP): two M = 128 tiles per pipeline stage.  Every
  output pixel sees the same MMAs in the same order as in the round-1 kernel, so the two must agree BIT FOR BIT — outputs,
  fused max-pool, fused 1x1 head and the masked data-gradient epilogue alike.
* PNNP_CONVT_FAST — ConvTranspose2d layers: compile-time specialised pixel-shuffle epilogue, weights resident in shared memory.
* PNNP_IN_V2 — NCHW fp32 -> NHWC16 bf16 input conversion, four pixels per thread.
* PNNP_CONV_F32X2 — the specialised 3x3 epilogues' fp32 arithmetic in packed pairs (FADD2 / FFMA2: the same IEEE operations).
* PNNP_WGRAD_V2 — weight-gradient kernel with the producer / MMA warps' loop-invariant state in registers.
* PNNP_E2E_ZERO_COPY=1 (still opt-in: measured slower, 9 854 against 10 314 MP/s) — HostSynthPipeline as ONE launch that reads /
  writes the pinned host buffers directly over PCIe.
* PNNP_COPY_V2 / PNNP_ACTBWD_V2 — 32-bit index arithmetic in the weight-packing copy / activation backward + bias gradient.
* PNNP_SSIM_V2 — separable 7x7 window sums in the eval epilogue (equal to float64 summation order).
* PNNP_CONV_PDL — conv layers launched with programmatic stream serialization (`griddepcontrol.wait` before the first global
  access)."""
import os

import pytest
import torch

import pnnp_b200 as P
from pnnp_b200 import _lib, archs

pytestmark = pytest.mark.gpu

VARIANT_SWITCHES = ("PNNP_CONV_SUPER", "PNNP_CONVT_FAST", "PNNP_IN_V2", "PNNP_CONV_PDL", "PNNP_CONV_F32X2", "PNNP_WGRAD_V2",
                    "PNNP_COPY_V2", "PNNP_ACTBWD_V2", "PNNP_SSIM_V2")


@pytest.fixture(autouse=True)
def _round1_kernels_as_baseline(monkeypatch):
    """Every test starts from the round-1 kernels (all switches 0) and turns on what it examines."""
    for name in VARIANT_SWITCHES:
        monkeypatch.setenv(name, "0")
    yield


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _layer(mode_name, cin, cout, two):
    g = torch.Generator(device="cuda").manual_seed(cin * 31 + cout)
    ct = cin * (2 if two else 1)
    wt = torch.randn((cout, ct, 3, 3), device="cuda", generator=g) / (3 * ct ** 0.5)
    b = torch.randn((cout,), device="cuda", generator=g) * 0.1

    class M:
        pass
    m = M()
    m.weight, m.bias = wt, None
    return archs._PackedLayer(m, mode_name).get(wt.device)[0], b, g


def _run(monkeypatch, sup, fn):
    if sup:
        monkeypatch.setenv("PNNP_CONV_SUPER", str(sup))
    else:
        monkeypatch.setenv("PNNP_CONV_SUPER", "0")
    out = fn()
    torch.cuda.synchronize()
    assert _lib.lib().pnnp_conv_pipeline_error() == 0, "tcgen05/TMA pipeline wait timed out"
    return out


@pytest.mark.parametrize("sup", [1, 2])
@pytest.mark.parametrize("xmode,cin,cout,h,w,n,two", [
    (True, 16, 32, 16, 32, 1, False), (True, 32, 32, 24, 44, 2, False), (True, 32, 32, 40, 30, 1, True),      # h % 16 = 8 edges
    (True, 32, 32, 512, 512, 2, False), (False, 32, 64, 32, 48, 1, False), (False, 64, 64, 72, 80, 2, False),
    (False, 64, 128, 24, 32, 1, False), (False, 128, 128, 48, 48, 1, False), (False, 64, 64, 256, 256, 1, True)])
def test_super_tile_equals_default_kernel_bit_for_bit(monkeypatch, sup, xmode, cin, cout, h, w, n, two):
    wp, b, g = _layer("conv3x" if xmode else "conv", cin, cout, two)
    mode = _lib.CONV3X if xmode else _lib.CONV3
    x = _nhwc(torch.randn((n, cin, h, w), device="cuda", generator=g))
    x2 = _nhwc(torch.randn((n, cin, h, w), device="cuda", generator=g)) if two else None

    def call():
        out = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device="cuda")
        pooled = torch.zeros((n, h // 2, w // 2, cout), dtype=torch.bfloat16, device="cuda")
        archs._conv(mode, x, wp, b, out, cout, _lib.ACT_LEAKY, x1=x2, pool_out=pooled)
        return out, pooled
    want = _run(monkeypatch, 0, call)
    got = _run(monkeypatch, sup, call)
    assert torch.equal(got[0].view(torch.int16), want[0].view(torch.int16))
    assert torch.equal(got[1].view(torch.int16), want[1].view(torch.int16))


@pytest.mark.parametrize("sup", [1, 2])
def test_super_tile_fused_head_and_mask(monkeypatch, sup):
    wp, b, g = _layer("conv", 32, 32, False)
    x = _nhwc(torch.randn((2, 32, 40, 48), device="cuda", generator=g))
    hw = torch.randn((4, 32), device="cuda", generator=g) / 6
    hb = torch.randn((4,), device="cuda", generator=g) * 0.1
    res = torch.randn((2, 4, 40, 48), device="cuda", generator=g)
    mask = _nhwc(torch.randn((2, 32, 40, 48), device="cuda", generator=g))

    def call():
        hout = torch.zeros((2, 4, 40, 48), dtype=torch.float32, device="cuda")
        archs._conv(_lib.CONV3, x, wp, b, None, 32, _lib.ACT_LEAKY, head=(hw, hb, hout), resid_nchw=res)
        dx = torch.zeros((2, 40, 48, 32), dtype=torch.bfloat16, device="cuda")
        archs._conv(_lib.CONV3, x, wp, None, dx, 32, _lib.ACT_NONE, mask=mask, mask_slope=0.2)
        return hout, dx
    want = _run(monkeypatch, 0, call)
    got = _run(monkeypatch, sup, call)
    assert torch.equal(got[0], want[0])
    assert torch.equal(got[1].view(torch.int16), want[1].view(torch.int16))


@pytest.mark.parametrize("sup", [1, 2])
def test_super_tile_whole_unet_forward_is_unchanged(monkeypatch, sup):
    torch.manual_seed(3)
    net = P.UNetSeeInDark({"in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": False}).cuda().eval()
    P.initialize_weights(net)
    x = torch.rand((1, 4, 208, 272), device="cuda")

    def call():
        with torch.no_grad():
            return net(x).clone()
    want = _run(monkeypatch, 0, call)
    got = _run(monkeypatch, sup, call)
    assert torch.equal(got, want)


def _run_env(monkeypatch, name, value, fn):
    if value:
        monkeypatch.setenv(name, str(value))
    else:
        monkeypatch.setenv(name, "0")
    out = fn()
    torch.cuda.synchronize()
    assert _lib.lib().pnnp_conv_pipeline_error() == 0, "tcgen05/TMA pipeline wait timed out"
    return out


@pytest.mark.parametrize("cin,cout,h,w,n", [(64, 32, 16, 32, 1), (64, 32, 356, 532, 1), (128, 64, 24, 40, 2), (256, 128, 8, 24, 1),
                                            (512, 256, 8, 8, 1)])
def test_conv_transpose_fast_path_equals_default(monkeypatch, cin, cout, h, w, n):
    g = torch.Generator(device="cuda").manual_seed(cin + h)
    x = _nhwc(torch.randn((n, cin, h, w), device="cuda", generator=g))
    wt = torch.randn((cin, cout, 2, 2), device="cuda", generator=g) / cin ** 0.5
    b = torch.randn((cout,), device="cuda", generator=g) * 0.1

    class M:
        pass
    m = M()
    m.weight, m.bias = wt, None
    wp = archs._PackedLayer(m, "convT").get(wt.device)[0]

    def call():
        out = torch.zeros((n, 2 * h, 2 * w, cout), dtype=torch.bfloat16, device="cuda")
        archs._conv(_lib.CONVT, x, wp, b, out, cout, _lib.ACT_NONE)
        return out
    want = _run_env(monkeypatch, "PNNP_CONVT_FAST", 0, call)
    got = _run_env(monkeypatch, "PNNP_CONVT_FAST", 1, call)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("shape", [(1, 4, 64, 96), (2, 4, 30, 34), (1, 4, 1424, 2128), (1, 12, 32, 48)])
def test_input_layout_v2_equals_default(monkeypatch, shape):
    x = torch.rand(shape, device="cuda") - 0.3

    def call():
        out = torch.zeros((shape[0], shape[2], shape[3], 16), dtype=torch.bfloat16, device="cuda")
        return archs._to_nhwc16(x, out, 0.75)
    want = _run_env(monkeypatch, "PNNP_IN_V2", 0, call)
    got = _run_env(monkeypatch, "PNNP_IN_V2", 1, call)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))


def test_all_variants_together_leave_the_unet_forward_unchanged(monkeypatch):
    torch.manual_seed(4)
    net = P.UNetSeeInDark({"in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": False}).cuda().eval()
    P.initialize_weights(net)
    x = torch.rand((1, 4, 304, 400), device="cuda")
    with torch.no_grad():
        want = net(x).clone()
        for k, v in (("PNNP_CONV_SUPER", "1"), ("PNNP_CONVT_FAST", "1"), ("PNNP_IN_V2", "1"), ("PNNP_CONV_PDL", "1")):
            monkeypatch.setenv(k, v)
        got = net(x).clone()
    torch.cuda.synchronize()
    assert _lib.lib().pnnp_conv_pipeline_error() == 0
    assert torch.equal(got, want)


@pytest.mark.parametrize("sup", [0, 1, 2])
def test_programmatic_dependent_launch_leaves_both_networks_unchanged(monkeypatch, sup):
    """Back-to-back conv layers under PDL: every layer must still see its predecessor's complete output (UNet and ResUnet,
    repeated forwards so that a layer of forward k + 1 follows the last layer of forward k)."""
    torch.manual_seed(6)
    arch = {"in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": False}
    for cls in (P.UNetSeeInDark, P.ResUnet):
        net = cls(arch).cuda().eval()
        P.initialize_weights(net)
        xs = [torch.rand((1, 4, 176, 240), device="cuda") for _ in range(4)]
        with torch.no_grad():
            want = [net(x).clone() for x in xs]
            monkeypatch.setenv("PNNP_CONV_PDL", "1")
            if sup:
                monkeypatch.setenv("PNNP_CONV_SUPER", str(sup))
            got = [net(x).clone() for x in xs]
            monkeypatch.setenv("PNNP_CONV_PDL", "0")
            monkeypatch.setenv("PNNP_CONV_SUPER", "0")
        torch.cuda.synchronize()
        assert _lib.lib().pnnp_conv_pipeline_error() == 0
        assert all(torch.equal(a, b) for a, b in zip(got, want))


@pytest.mark.parametrize("combo", [{"PNNP_CONV_F32X2": "1"},
                                   {"PNNP_CONV_F32X2": "1", "PNNP_CONV_SUPER": "1", "PNNP_CONV_PDL": "1"},
                                   {"PNNP_CONV_F32X2": "1", "PNNP_CONV_SUPER": "2", "PNNP_CONV_PDL": "1"}])
def test_packed_pair_epilogue_is_bit_identical(monkeypatch, combo):
    """x-mode combine, bias + LeakyReLU, fused pool and fused head through FADD2 / FFMA2, per layer and for both networks."""
    wp, b, g = _layer("conv3x", 32, 32, False)
    x = _nhwc(torch.randn((2, 32, 40, 44), device="cuda", generator=g))
    hw = torch.randn((4, 32), device="cuda", generator=g) / 6
    hb = torch.randn((4,), device="cuda", generator=g) * 0.1
    wp2, b2, _ = _layer("conv", 64, 64, False)
    x2 = _nhwc(torch.randn((1, 64, 24, 48), device="cuda", generator=g))
    torch.manual_seed(11)
    nets = []
    for cls in (P.UNetSeeInDark, P.ResUnet):
        net = cls({"in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": False}).cuda().eval()
        P.initialize_weights(net)
        nets.append(net)
    frame = torch.rand((1, 4, 144, 208), device="cuda")

    def call():
        out = torch.zeros((2, 40, 44, 32), dtype=torch.bfloat16, device="cuda")
        pooled = torch.zeros((2, 20, 22, 32), dtype=torch.bfloat16, device="cuda")
        archs._conv(_lib.CONV3X, x, wp, b, out, 32, _lib.ACT_LEAKY, pool_out=pooled)
        hout = torch.zeros((2, 4, 40, 44), dtype=torch.float32, device="cuda")
        archs._conv(_lib.CONV3X, x, wp, b, None, 32, _lib.ACT_LEAKY, head=(hw, hb, hout))
        out2 = torch.zeros((1, 24, 48, 64), dtype=torch.bfloat16, device="cuda")
        archs._conv(_lib.CONV3, x2, wp2, b2, out2, 64, _lib.ACT_RELU)
        with torch.no_grad():
            full = [net(frame).clone() for net in nets]
        return [out.view(torch.int16), pooled.view(torch.int16), hout, out2.view(torch.int16)] + full
    for k in ("PNNP_CONV_F32X2", "PNNP_CONV_SUPER", "PNNP_CONV_PDL"):
        monkeypatch.setenv(k, "0")
    want = call()
    for k, v in combo.items():
        monkeypatch.setenv(k, v)
    got = call()
    torch.cuda.synchronize()
    assert _lib.lib().pnnp_conv_pipeline_error() == 0
    assert all(torch.equal(a, c) for a, c in zip(got, want))


@pytest.mark.parametrize("shape,scale,correct", [((1, 4, 64, 96), 1.0, False), ((2, 3, 45, 70), 1.5, False), ((1, 4, 128, 192), 1.0, True)])
def test_separable_ssim_equals_default_kernel(monkeypatch, shape, scale, correct):
    from pnnp_b200.metrics import eval_partial_sums
    g = torch.Generator(device="cuda").manual_seed(9)
    hr = torch.rand(shape, device="cuda", generator=g)
    dn = ((hr + 0.07 * torch.randn(shape, device="cuda", generator=g)) / scale).contiguous()
    monkeypatch.setenv("PNNP_SSIM_V2", "0")
    want = eval_partial_sums(dn, hr, scale, correct).clone()
    monkeypatch.setenv("PNNP_SSIM_V2", "1")
    got = eval_partial_sums(dn, hr, scale, correct)
    assert torch.allclose(got, want, rtol=1e-11, atol=1e-9), (got, want)


def test_wgrad_v2_matches_autograd_like_the_default_kernel(monkeypatch):
    """The register-resident producer / MMA loops of wgrad_nhwc_kernel<1> through the default kernel's own tests: every layer
    shape (all operand modes: filter rows in M, tap-in-N, per-column CTAs, two sources, transposed conv) and a whole step."""
    import test_gpu_train as T
    monkeypatch.setenv("PNNP_WGRAD_V2", "1")
    for args in [(16, 32, 16, 32, 1), (32, 32, 24, 40, 2), (32, 64, 20, 36, 1), (64, 64, 16, 48, 2), (64, 128, 16, 16, 2),
                 (128, 64, 16, 32, 1), (128, 128, 24, 16, 1), (256, 256, 8, 16, 1), (512, 256, 8, 8, 1), (256, 512, 4, 6, 2)]:
        T.test_wgrad_nhwc_3x3_matches_autograd(*args)
    T.test_wgrad_nhwc_two_sources_accumulate_into_one_gradient()
    for args in [(64, 32, 16, 32), (128, 64, 8, 24), (512, 256, 8, 8)]:
        T.test_wgrad_nhwc_conv_transpose(*args)
    T.test_training_step_gradients_match_fp32_autograd(None)
    torch.cuda.synchronize()
    assert _lib.lib().pnnp_wgrad_nhwc_pipeline_error() == 0


def test_zero_copy_host_pipeline_equals_the_chunked_pipeline():
    import numpy as np
    from pnnp_b200.pipeline import HostSynthPipeline
    n, c, h, w = 12, 4, 64, 128
    np.random.seed(3)
    params = [P.sample_params("SonyA7S2") for _ in range(n)]
    host_in = (torch.rand((n, c, h, w)) ** 2).pin_memory()
    outs = []
    for zc in (False, True):
        pipe = HostSynthPipeline(n, c, h, w, torch.device("cuda", 0), chunk=5, n_streams=2, zero_copy=zc)
        host_out = torch.empty_like(host_in).pin_memory()
        pipe.run(host_in, host_out, params, "pgrq", generator=P.PhiloxGenerator(21), crop_id0=7, post_clip=(-float("inf"), 1.0))
        torch.cuda.synchronize()
        outs.append(host_out.clone())
    assert torch.equal(outs[0], outs[1]) and float(outs[1].max()) <= 1.0 and not torch.equal(outs[1], host_in)


def test_copy_v2_training_step_is_unchanged(monkeypatch):
    """Weight packing and gradient re-layout through the 32-bit-index copy kernel: forward output bit-identical (same packed
    weights), one training step equal to the default's up to the fp32 atomics of the weight-gradient kernel."""
    from pnnp_b200.train import UNetTrainStep
    res = {}
    for v2 in (0, 1):
        if v2:
            monkeypatch.setenv("PNNP_COPY_V2", "1")
        else:
            monkeypatch.setenv("PNNP_COPY_V2", "0")
        torch.manual_seed(13)
        net = P.UNetSeeInDark({"in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": False}).cuda()
        P.initialize_weights(net)
        g = torch.Generator(device="cuda").manual_seed(2)
        hr = torch.rand((2, 4, 64, 64), device="cuda", generator=g)
        lr = (hr + 0.05 * torch.randn((2, 4, 64, 64), device="cuda", generator=g)).contiguous()
        ts = UNetTrainStep(net.train(), lr=1e-3)
        ts.use_graph = False
        loss = float(ts.step(lr, hr))
        res[v2] = (loss, ts.scr.bufs["pred"].clone(), ts.flat_p.clone())
    assert torch.equal(res[0][1], res[1][1])                               # prediction of the first step: packed weights are the same bits
    assert abs(res[0][0] - res[1][0]) < 1e-7
    assert (res[0][2] - res[1][2]).abs().max().item() < 2.5e-3            # Adam's first step is sign-like: |update| <= lr


@pytest.mark.parametrize("act_kind", [0, 1, 2])
def test_act_backward_v2_equals_the_default_kernel(monkeypatch, act_kind):
    """pnnp_act_bwd_bias through the second kernel form: the in-place gradient is the same bits, the bias gradient the same sums
    (fp32 atomics: order differs), on the channel counts of the UNet (c / 8 a power of two) and a ragged pixel count."""
    L = _lib.lib()
    for pixels, c in ((4099, 16), (64 * 64 * 2, 32), (3001, 64), (1025, 256), (300, 512)):
        g0 = torch.randn((pixels, c), device="cuda").to(torch.bfloat16)
        out = torch.randn((pixels, c), device="cuda").to(torch.bfloat16)
        res = {}
        for v2 in (0, 1):
            if v2:
                monkeypatch.setenv("PNNP_ACTBWD_V2", "1")
            else:
                monkeypatch.setenv("PNNP_ACTBWD_V2", "0")
            g = g0.clone()
            db = torch.full((c,), 0.25, device="cuda")
            _lib.check(L.pnnp_act_bwd_bias(g.data_ptr(), out.data_ptr(), db.data_ptr(), pixels, c, act_kind, _lib.stream_ptr(g.device)),
                       "act_bwd_bias")
            torch.cuda.synchronize()
            res[v2] = (g, db)
        assert torch.equal(res[0][0], res[1][0]), (pixels, c)
        assert torch.allclose(res[0][1], res[1][1], rtol=1e-4, atol=1e-3), (pixels, c)
