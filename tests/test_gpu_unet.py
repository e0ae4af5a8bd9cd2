"""tcgen05 implicit-GEMM layers and the UNet forward through the C ABI vs torch fp32 / the oracle.

Tolerances (stated, per BASELINE north_star): the network output must be within 1e-3 max-abs of
the fp32 reference with the reference's own init; this build computes in bf16 with fp32
accumulation (the "bf16 variant"), so per-layer checks use a bf16-appropriate relative bound
against an fp32 reference evaluated on the SAME bf16-rounded inputs and weights."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle_np as O
import pnnp_b200 as P
from pnnp_b200 import _lib, archs

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).float()


def _nhwc(t):       # NCHW fp32 -> NHWC bf16
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(t):       # NHWC bf16 -> NCHW fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def _pack(w, kind="conv"):
    class M:        # minimal stand-in with .weight/.bias for _PackedLayer
        pass
    m = M()
    m.weight, m.bias = w, None
    return archs._PackedLayer(m, kind).get(w.device)[0]


def _no_pipeline_error():
    torch.cuda.synchronize()
    assert _lib.lib().pnnp_conv_pipeline_error() == 0, "tcgen05/TMA pipeline wait timed out"
    assert _lib.lib().pnnp_conv_first_pipeline_error() == 0, "first-layer kernel: MMA completion wait timed out"


@pytest.mark.parametrize("cin,cout,h,w,n", [(16, 32, 16, 32, 1), (32, 32, 24, 40, 2), (64, 64, 32, 48, 1),
                                            (128, 256, 16, 16, 1), (256, 512, 8, 16, 1), (512, 512, 16, 32, 1),
                                            (64, 32, 40, 72, 1)])
@pytest.mark.parametrize("act", [_lib.ACT_LEAKY, _lib.ACT_NONE])
def test_conv3x3_layer(cin, cout, h, w, n, act):
    g = torch.Generator(device="cuda").manual_seed(cin * 1000 + cout)
    x = torch.randn((n, cin, h, w), device="cuda", generator=g)
    wt = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) / (3 * cin ** 0.5)
    b = torch.randn((cout,), device="cuda", generator=g) * 0.1
    out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device="cuda")
    archs._conv(_lib.CONV3, _nhwc(x), _pack(wt), b, out, cout, act)
    _no_pipeline_error()
    ref = F.conv2d(_bf(x), _bf(wt), b, padding=1)
    if act == _lib.ACT_LEAKY:
        ref = F.leaky_relu(ref, 0.2)
    err = (_nchw(out) - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err      # bf16 output rounding: 2^-8 relative


def test_conv3x3_two_sources_is_concat():
    g = torch.Generator(device="cuda").manual_seed(5)
    up, skip = torch.randn((1, 64, 24, 32), device="cuda", generator=g), torch.randn((1, 64, 24, 32), device="cuda", generator=g)
    wt = torch.randn((64, 128, 3, 3), device="cuda", generator=g) / 30
    b = torch.randn((64,), device="cuda", generator=g) * 0.1
    out = torch.empty((1, 24, 32, 64), dtype=torch.bfloat16, device="cuda")
    archs._conv(_lib.CONV3, _nhwc(up), _pack(wt), b, out, 64, _lib.ACT_LEAKY, x1=_nhwc(skip))
    _no_pipeline_error()
    ref = F.leaky_relu(F.conv2d(torch.cat([_bf(up), _bf(skip)], 1), _bf(wt), b, padding=1), 0.2)
    assert (_nchw(out) - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("cin,cout", [(512, 256), (64, 32)])
def test_conv_transpose_layer(cin, cout):
    g = torch.Generator(device="cuda").manual_seed(cin)
    x = torch.randn((1, cin, 8, 24), device="cuda", generator=g)
    wt = torch.randn((cin, cout, 2, 2), device="cuda", generator=g) / cin ** 0.5
    b = torch.randn((cout,), device="cuda", generator=g) * 0.1
    out = torch.empty((1, 16, 48, cout), dtype=torch.bfloat16, device="cuda")
    archs._conv(_lib.CONVT, _nhwc(x), _pack(wt, "convT"), b, out, cout, _lib.ACT_NONE)
    _no_pipeline_error()
    ref = F.conv_transpose2d(_bf(x), _bf(wt), b, stride=2)
    assert (_nchw(out) - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


def test_conv1x1_to_nchw_f32_with_residual():
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn((2, 32, 24, 40), device="cuda", generator=g)
    wt = torch.randn((4, 32, 1, 1), device="cuda", generator=g) / 6
    b = torch.randn((4,), device="cuda", generator=g) * 0.1
    res = torch.randn((2, 4, 24, 40), device="cuda", generator=g)
    out = torch.empty((2, 4, 24, 40), dtype=torch.float32, device="cuda")
    archs._conv(_lib.CONV1, _nhwc(x), _pack(wt), b, out, 4, _lib.ACT_NONE, out_mode=_lib.OUT_NCHW_F32, resid_nchw=res)
    _no_pipeline_error()
    ref = F.conv2d(_bf(x), _bf(wt), b) + res
    assert (out - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())   # fp32 accumulators, fp32 output


def test_maxpool_and_input_layout():
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((2, 64, 16, 48), device="cuda", generator=g)
    out = torch.empty((2, 8, 24, 64), dtype=torch.bfloat16, device="cuda")
    archs._pool(_nhwc(x), out)
    assert torch.equal(_nchw(out), F.max_pool2d(_bf(x), 2))
    x4 = torch.rand((2, 4, 16, 32), device="cuda", generator=g)
    o16 = torch.empty((2, 16, 32, 16), dtype=torch.bfloat16, device="cuda")
    archs._to_nhwc16(x4, o16)
    assert torch.equal(o16[..., :4].float(), _bf(x4).permute(0, 2, 3, 1)) and float(o16[..., 4:].abs().max()) == 0.0


def _arch(res=False):
    return dict(name="UNetSeeInDark", in_nc=4, out_nc=4, nf=32, nframes=1, use_dpsv=False, res=res, cascade=False,
                add=False, lock_wb=False)


@pytest.mark.parametrize("res", [False, True])
@pytest.mark.parametrize("shape", [(1, 4, 64, 96), (2, 4, 512, 512)])
def test_unet_forward_vs_fp32_oracle_reference_init(shape, res):
    """Reference init (N(0, 0.02)): |out - fp32 oracle| <= 1e-3 max-abs (north_star) and PSNR drift
    far below 0.01 dB on [0,1] data."""
    torch.manual_seed(7)
    net = P.UNetSeeInDark(_arch(res)).cuda()
    P.initialize_weights(net)
    net.eval()
    x = torch.rand(shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1997))
    with torch.no_grad():
        got = net(x)
        want = O.unet_forward(x, net.state_dict(), res=res)       # torch fp32 functional restatement (cuDNN fp32)
    _no_pipeline_error()
    err = (got - want).abs().max().item()
    rel = err / want.abs().max().item()
    print(f"unet {shape} res={res}: max-abs {err:.3e}, rel {rel:.3e}, out absmax {want.abs().max().item():.3e}")
    assert err <= 1e-3, err
    target = torch.rand_like(want)
    psnr = lambda a: 10 * torch.log10(1.0 / ((a.clamp(0, 1) - target) ** 2).mean())
    assert abs(psnr(got).item() - psnr(want).item()) < 0.01


def test_unet_forward_default_init_relative_error():
    """PyTorch default init (outputs O(0.1)): report bf16-vs-fp32 relative error (bf16 variant)."""
    torch.manual_seed(3)
    net = P.UNetSeeInDark(_arch()).cuda().eval()
    x = torch.rand((1, 4, 128, 160), device="cuda")
    with torch.no_grad():
        got, want = net(x), O.unet_forward(x, net.state_dict())
    _no_pipeline_error()
    rel = ((got - want).abs().max() / want.abs().max()).item()
    print(f"default init: rel max err {rel:.3e}")
    assert rel < 3e-2


def test_state_dict_is_reference_compatible(golden):
    """Keys / shapes equal the reference module's (golden fixture made from the live reference)."""
    g = golden("nets")
    ref_keys = [k.split("__sd__")[1] for k in g.files if k.startswith("UNetSeeInDark_res0__sd__")]
    arch = _arch()
    arch["nf"] = 16
    net = P.UNetSeeInDark(arch)
    assert list(net.state_dict().keys()) == ref_keys


def test_rejects_bad_shapes():
    net = P.UNetSeeInDark(_arch()).cuda().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="multiples of 16"):
        net(torch.rand((1, 4, 40, 64), device="cuda"))


@pytest.mark.parametrize("cin,cout,h,w", [(32, 64, 32, 64), (64, 128, 16, 32), (256, 512, 32, 32), (32, 64, 24, 40)])
def test_conv3x3_stride2_layer(cin, cout, h, w):
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    x = torch.randn((2, cin, h, w), device="cuda", generator=g)
    wt = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) / (3 * cin ** 0.5)
    b = torch.randn((cout,), device="cuda", generator=g) * 0.1
    out = torch.empty((2, h // 2, w // 2, cout), dtype=torch.bfloat16, device="cuda")
    archs._conv(_lib.CONV3S2, _nhwc(x), _pack(wt), b, out, cout, _lib.ACT_NONE)
    _no_pipeline_error()
    ref = F.conv2d(_bf(x), _bf(wt), b, padding=1, stride=2)
    assert (_nchw(out) - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


def test_residual_add_in_epilogue():
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((1, 64, 16, 32), device="cuda", generator=g)
    wt = torch.randn((64, 64, 3, 3), device="cuda", generator=g) / 24
    out = torch.empty((1, 16, 32, 64), dtype=torch.bfloat16, device="cuda")
    xb = _nhwc(x)
    archs._conv(_lib.CONV3, xb, _pack(wt), None, out, 64, _lib.ACT_NONE, resid=xb)
    _no_pipeline_error()
    ref = F.conv2d(_bf(x), _bf(wt), None, padding=1) + _bf(x)
    assert (_nchw(out) - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("shape", [(1, 4, 64, 96), (1, 4, 256, 512)])
def test_resunet_forward_vs_fp32_oracle_reference_init(shape):
    torch.manual_seed(11)
    net = P.ResUnet(_arch()).cuda()
    P.initialize_weights(net)
    net.eval()
    x = torch.rand(shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    with torch.no_grad():
        got = net(x)
        want = O.resunet_forward(x, net.state_dict())
    _no_pipeline_error()
    err = (got - want).abs().max().item()
    print(f"resunet {shape}: max-abs {err:.3e}, rel {err / want.abs().max().item():.3e}")
    assert err <= 1e-3, err


def test_resunet_state_dict_is_reference_compatible(golden):
    g = golden("nets")
    ref_keys = [k.split("__sd__")[1] for k in g.files if k.startswith("ResUnet_res0__sd__")]
    arch = _arch()
    arch["nf"] = 16
    net = P.ResUnet(arch)
    assert list(net.state_dict().keys()) == ref_keys
    assert sum(p.numel() for p in P.ResUnet(_arch()).parameters()) == 11075268


def test_fused_pool_and_head_epilogues():
    """conv + LeakyReLU + MaxPool2d(2) and conv + LeakyReLU + 1x1 head (+ residual) fused in the epilogue."""
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn((2, 32, 24, 48), device="cuda", generator=g)
    wt = torch.randn((32, 32, 3, 3), device="cuda", generator=g) / 17
    b = torch.randn((32,), device="cuda", generator=g) * 0.1
    out = torch.empty((2, 24, 48, 32), dtype=torch.bfloat16, device="cuda")
    pooled = torch.empty((2, 12, 24, 32), dtype=torch.bfloat16, device="cuda")
    archs._conv(_lib.CONV3, _nhwc(x), _pack(wt), b, out, 32, _lib.ACT_LEAKY, pool_out=pooled)
    _no_pipeline_error()
    assert torch.equal(_nchw(pooled), F.max_pool2d(_nchw(out), 2))
    hw = torch.randn((4, 32), device="cuda", generator=g) / 6
    hb = torch.randn((4,), device="cuda", generator=g) * 0.1
    res = torch.randn((2, 4, 24, 48), device="cuda", generator=g)
    hout = torch.empty((2, 4, 24, 48), dtype=torch.float32, device="cuda")
    archs._conv(_lib.CONV3, _nhwc(x), _pack(wt), b, None, 32, _lib.ACT_LEAKY, head=(hw, hb, hout), resid_nchw=res)
    _no_pipeline_error()
    act = F.leaky_relu(F.conv2d(_bf(x), _bf(wt), b, padding=1), 0.2)
    ref = F.conv2d(act, hw.view(4, 32, 1, 1), hb) + res
    assert (hout - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("cin,cout,h,w,n,two", [(16, 32, 16, 32, 1, False), (32, 32, 24, 44, 2, False), (64, 64, 40, 72, 1, False),
                                               (32, 32, 16, 30, 1, True), (64, 64, 24, 28, 1, True), (32, 64, 8, 16, 1, False)])
def test_conv3x3_x_shift_in_n_mode(cin, cout, h, w, n, two):
    """MODE_CONV3X == F.conv2d, including tiles_x = ceil(W/14) edges, two-source concat and the fused pool."""
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout + w)
    x = torch.randn((n, cin, h, w), device="cuda", generator=g)
    x2 = torch.randn((n, cin, h, w), device="cuda", generator=g) if two else None
    ct = cin * (2 if two else 1)
    wt = torch.randn((cout, ct, 3, 3), device="cuda", generator=g) / (3 * ct ** 0.5)
    b = torch.randn((cout,), device="cuda", generator=g) * 0.1

    class M:
        pass
    m = M()
    m.weight, m.bias = wt, None
    wp = archs._PackedLayer(m, "conv3x").get(wt.device)[0]
    out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device="cuda")
    pooled = torch.empty((n, h // 2, w // 2, cout), dtype=torch.bfloat16, device="cuda")
    archs._conv(_lib.CONV3X, _nhwc(x), wp, b, out, cout, _lib.ACT_LEAKY, x1=None if x2 is None else _nhwc(x2), pool_out=pooled)
    _no_pipeline_error()
    xin = _bf(x) if x2 is None else torch.cat([_bf(x), _bf(x2)], 1)
    ref = F.leaky_relu(F.conv2d(xin, _bf(wt), b, padding=1), 0.2)
    assert (_nchw(out) - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert torch.equal(_nchw(pooled), F.max_pool2d(_nchw(out), 2))


def _cpu_state(net):
    return {k: v.detach().cpu() for k, v in net.state_dict().items()}


@pytest.mark.parametrize("arch_name", ["UNetSeeInDark", "ResUnet"])
@pytest.mark.parametrize("shape,reflect", [((1, 4, 1424, 2128), False),        # BASELINE configs[0] / [4]: Sony frame, 89 x 133 bottleneck
                                           ((1, 4, 1744, 2320), False),        # configs[3]: the padded IMX686 extent the network sees
                                           ((1, 4, 1736, 2312), True)])        # configs[3] through the reflect-pad branch
def test_full_frame_forward_vs_cpu_fp32_oracle(arch_name, shape, reflect):
    """U1 / U2 at the BASELINE frame shapes against the oracle's fp32 forward run on the CPU (independent of cuDNN): max-abs
    <= 1e-3 with the reference's initialiser (north_star), relative error printed.  `reflect`: the eval boundary of
    trainer_LRID.py:224-229 / trainer_SID.py:221-226 — F.pad(reflect, 4) -> net -> crop 4 — as the product's trainer runs it
    (2312 % 16 == 8); the oracle pads, runs and crops on the CPU."""
    from pnnp_b200.trainer import SID_Trainer
    torch.manual_seed(23)
    net = getattr(P, arch_name)(_arch()).cuda()
    P.initialize_weights(net)
    net.eval()
    x = torch.rand(shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1997))
    oracle_fwd = O.unet_forward if arch_name == "UNetSeeInDark" else O.resunet_forward
    with torch.no_grad():
        if reflect:
            holder = type("T", (), {"net": net})()
            got = SID_Trainer.forward_frame(holder, x)
            want = oracle_fwd(F.pad(x.cpu(), (4, 4, 4, 4), mode="reflect"), _cpu_state(net))[..., 4:-4, 4:-4]
        else:
            got = net(x)
            want = oracle_fwd(x.cpu(), _cpu_state(net))
    _no_pipeline_error()
    assert got.shape == x.shape
    err = (got.cpu() - want).abs().max().item()
    rel = err / want.abs().max().item()
    print(f"{arch_name} {shape} reflect={reflect}: max-abs {err:.3e}, rel {rel:.3e}, out absmax {want.abs().max().item():.3e}")
    assert err <= 1e-3, err
    target = torch.rand(want.shape, generator=torch.Generator().manual_seed(3))
    psnr = lambda a: 10 * torch.log10(1.0 / ((a.clamp(0, 1) - target) ** 2).mean())
    assert abs(psnr(got.cpu()).item() - psnr(want).item()) < 0.01           # PSNR +- 0.01 dB on the same inputs


@pytest.mark.parametrize("cin,cout,h,w,n,act", [(4, 32, 16, 32, 1, _lib.ACT_LEAKY), (4, 32, 21, 45, 2, _lib.ACT_RELU), (3, 16, 9, 17, 1, _lib.ACT_NONE),
                                                (4, 64, 24, 16, 1, _lib.ACT_LEAKY), (1, 48, 8, 40, 1, _lib.ACT_LEAKY),
                                                (4, 32, 512, 512, 3, _lib.ACT_LEAKY), (4, 32, 1424, 2128, 1, _lib.ACT_LEAKY)])
def test_fused_first_layer(cin, cout, h, w, n, act):
    """csrc/conv_first.cu — packed NCHW fp32 planes -> conv3x3 + bias + activation -> NHWC bf16 in one launch (the pack boundary
    fused into the first layer: process.py:625-631 -> Unet.py:55) — against torch on the bf16-rounded operands, ragged tile edges
    and the BASELINE crop / frame shapes included."""
    g = torch.Generator(device="cuda").manual_seed(cin * 100 + cout + h)
    x = torch.randn((n, cin, h, w), device="cuda", generator=g)
    m = torch.nn.Conv2d(cin, cout, 3, padding=1).cuda()
    with torch.no_grad():
        m.weight.copy_(torch.randn((cout, cin, 3, 3), device="cuda", generator=g) / (3 * cin ** 0.5))
        m.bias.copy_(torch.randn((cout,), device="cuda", generator=g) * 0.1)
    out = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    archs._first_conv(x, m, out, act)
    _no_pipeline_error()
    ref = F.conv2d(_bf(x), _bf(m.weight.detach()), m.bias.detach(), padding=1)
    ref = F.leaky_relu(ref, 0.2) if act == _lib.ACT_LEAKY else (F.relu(ref) if act == _lib.ACT_RELU else ref)
    assert not torch.isnan(out.float()).any()
    assert (_nchw(out) - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


def test_fused_first_layer_warp_specialised_equals_single_role_kernel(monkeypatch):
    """conv_first_ws_kernel (producer / MMA / epilogue warps, double-buffered operands and accumulators) is bit-identical to the
    single-role conv_first_kernel: same operand layouts, same MMAs, same epilogue arithmetic."""
    g = torch.Generator(device="cuda").manual_seed(77)
    for cin, cout, h, w, n in ((4, 32, 40, 52, 2), (3, 64, 17, 33, 1), (4, 16, 8, 16, 3), (4, 32, 512, 512, 5), (4, 32, 1424, 2128, 1)):
        x = torch.randn((n, cin, h, w), device="cuda", generator=g)
        m = torch.nn.Conv2d(cin, cout, 3, padding=1).cuda()
        outs = []
        for ws in ("0", "1"):
            monkeypatch.setenv("PNNP_FIRST_WS", ws)
            out = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
            archs._first_conv(x, m, out, _lib.ACT_LEAKY)
            outs.append(out)
        monkeypatch.delenv("PNNP_FIRST_WS")
        _no_pipeline_error()
        assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16)), (cin, cout, h, w, n)


def test_fused_first_layer_equals_the_two_kernel_path(monkeypatch):
    """Same operands, same products: the fused first layer and input conversion + general conv kernel agree to the accumulation
    order of 36 fp32 terms (bf16 outputs equal except for isolated last-bit roundings), and so do the network outputs."""
    torch.manual_seed(8)
    x = torch.rand((2, 4, 208, 272), device="cuda")
    for cls in (P.UNetSeeInDark, P.ResUnet):
        net = cls(_arch()).cuda().eval()
        P.initialize_weights(net)
        with torch.no_grad():
            monkeypatch.setenv("PNNP_FUSED_FIRST", "0")
            want = net(x).clone()
            monkeypatch.setenv("PNNP_FUSED_FIRST", "1")
            got = net(x).clone()
        _no_pipeline_error()
        assert (got - want).abs().max().item() < 2e-3 * want.abs().max().item(), cls.__name__      # bf16 rounding flips of the first layer, propagated


@pytest.mark.parametrize("cin,cout,h,w,two,xmode", [(16, 32, 16, 32, False, False), (32, 64, 24, 40, False, False), (128, 128, 16, 16, False, False),
                                                    (256, 512, 8, 16, False, False), (64, 64, 24, 32, True, False), (32, 32, 24, 44, False, True),
                                                    (32, 32, 16, 30, True, True)])
def test_conv3x3_layer_tf32_variant(cin, cout, h, w, two, xmode):
    """The fp32-storage variant (tcgen05 kind::tf32, fp32 NHWC in and out): against an fp32 reference the error is that of 10-bit
    mantissa products — far below the bf16 variant's."""
    g = torch.Generator(device="cuda").manual_seed(cin * 13 + cout)
    x = torch.randn((1, cin, h, w), device="cuda", generator=g)
    x2 = torch.randn((1, cin, h, w), device="cuda", generator=g) if two else None
    ct = cin * (2 if two else 1)
    wt = torch.randn((cout, ct, 3, 3), device="cuda", generator=g) / (3 * ct ** 0.5)
    b = torch.randn((cout,), device="cuda", generator=g) * 0.1

    class M:
        pass
    m = M()
    m.weight, m.bias = wt, None
    wp = archs._PackedLayer(m, "conv3x" if xmode else "conv", torch.float32).get(wt.device)[0]
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    out = torch.full((1, h, w, cout), float("nan"), dtype=torch.float32, device="cuda")
    pooled = torch.empty((1, h // 2, w // 2, cout), dtype=torch.float32, device="cuda")
    archs._conv(_lib.CONV3X if xmode else _lib.CONV3, nhwc(x), wp, b, out, cout, _lib.ACT_LEAKY, x1=None if x2 is None else nhwc(x2),
                pool_out=pooled)
    _no_pipeline_error()
    xin = x if x2 is None else torch.cat([x, x2], 1)
    ref = F.leaky_relu(F.conv2d(xin, wt, b, padding=1), 0.2)
    got = out.permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err                       # tf32: 2^-11 relative per product
    assert torch.equal(pooled.permute(0, 3, 1, 2), F.max_pool2d(got, 2))


@pytest.mark.parametrize("arch_name", ["UNetSeeInDark", "ResUnet"])
def test_forward_tf32_variant_meets_the_fp32_bound_under_any_init(arch_name):
    """north_star: "UNet output within 1e-3 max-abs in fp32 (bf16 variant reported separately)".  With PyTorch's DEFAULT init the
    outputs are O(0.1-1) — there the bf16 variant's absolute error grows with them, the fp32-storage / tf32 variant stays inside
    1e-3; with the reference's initialiser both do.  Both variants are reported (printed) against the CPU fp32 oracle."""
    oracle_fwd = O.unet_forward if arch_name == "UNetSeeInDark" else O.resunet_forward
    x = torch.rand((1, 4, 256, 384), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    for init in ("default", "reference"):
        torch.manual_seed(3)
        net = getattr(P, arch_name)(_arch()).cuda().eval()
        if init == "reference":
            P.initialize_weights(net)
        with torch.no_grad():
            want = oracle_fwd(x.cpu(), _cpu_state(net))
            net.precision = "bf16"
            e_bf16 = (net(x).cpu() - want).abs().max().item()
            net.precision = "tf32"
            e_tf32 = (net(x).cpu() - want).abs().max().item()
        _no_pipeline_error()
        print(f"{arch_name} init={init}: out absmax {want.abs().max().item():.3e}; max-abs error bf16 {e_bf16:.3e}, tf32 {e_tf32:.3e}")
        assert e_tf32 <= 1e-3 and e_tf32 < e_bf16, (init, e_tf32, e_bf16)


@pytest.mark.parametrize("cin,cout,h,w,n", [(32, 32, 16, 8, 1), (32, 32, 24, 44, 2), (32, 32, 19, 29, 1), (32, 64, 40, 72, 1), (16, 32, 34, 18, 1),
                                            (64, 32, 20, 36, 1)])
def test_conv3x3_one_box_mode(cin, cout, h, w, n):
    """MODE_CONV3B: all nine taps from ONE haloed TMA box (8 x 16 tiles; tap = descriptor start offset of (dy * 10 + dx) pixel rows, SBO =
    the box's row pitch — tools/ubench_umma_offset.cu shows tcgen05.mma reading exactly those rows).  Bit-equal to the per-tap mode (same
    products accumulated in the same tap / channel order), for ragged sizes, 32 / 64 / 128-byte rows, and with the pool, head, residual
    and head + residual epilogues."""
    g = torch.Generator(device="cuda").manual_seed(cin * 1000 + cout + h)
    x = torch.randn((n, cin, h, w), device="cuda", generator=g)
    wt = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) / (3 * cin ** 0.5)
    b = torch.randn((cout,), device="cuda", generator=g) * 0.1
    xs, wp = _nhwc(x), _pack(wt)
    outs = {}
    for mode in (_lib.CONV3, _lib.CONV3B):
        o = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device="cuda")
        archs._conv(mode, xs, wp, b, o, cout, _lib.ACT_LEAKY)
        _no_pipeline_error()
        outs[mode] = o
    ref = F.leaky_relu(F.conv2d(_bf(x), _bf(wt), b, padding=1), 0.2)
    assert (_nchw(outs[_lib.CONV3B]) - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert torch.equal(outs[_lib.CONV3], outs[_lib.CONV3B])
    # residual epilogue
    r = torch.randn((n, h, w, cout), device="cuda", generator=g).to(torch.bfloat16)
    o3, ob = torch.empty_like(r), torch.empty_like(r)
    archs._conv(_lib.CONV3, xs, wp, b, o3, cout, _lib.ACT_NONE, resid=r)
    archs._conv(_lib.CONV3B, xs, wp, b, ob, cout, _lib.ACT_NONE, resid=r)
    _no_pipeline_error()
    assert torch.equal(o3, ob)
    if h % 2 == 0 and w % 2 == 0:                                   # fused 2x2 max-pool
        pooled = torch.empty((n, h // 2, w // 2, cout), dtype=torch.bfloat16, device="cuda")
        o = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device="cuda")
        archs._conv(_lib.CONV3B, xs, wp, b, o, cout, _lib.ACT_LEAKY, pool_out=pooled)
        _no_pipeline_error()
        assert torch.equal(o, outs[_lib.CONV3]) and torch.equal(_nchw(pooled), F.max_pool2d(_nchw(o), 2))
    # fused 1x1 head, with and without the residual
    hw = torch.randn((4, cout), device="cuda", generator=g) / 6
    hb = torch.randn((4,), device="cuda", generator=g) * 0.1
    res = torch.randn((n, 4, h, w), device="cuda", generator=g)
    for resid in (None, r):
        h3 = torch.empty((n, 4, h, w), dtype=torch.float32, device="cuda")
        hbm = torch.empty_like(h3)
        kw = dict(head=(hw, hb, h3), resid_nchw=res)
        if resid is not None:
            kw["resid"] = resid
        archs._conv(_lib.CONV3, xs, wp, b, None, cout, _lib.ACT_LEAKY, **kw)
        kw["head"] = (hw, hb, hbm)
        archs._conv(_lib.CONV3B, xs, wp, b, None, cout, _lib.ACT_LEAKY, **kw)
        _no_pipeline_error()
        assert torch.equal(h3, hbm)
