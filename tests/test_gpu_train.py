"""T1 — synthetic-pair training step (trainer_SID.py:93-101, losses/base_loss.py:92-103) through the C ABI.

Every backward kernel is checked against torch autograd in fp32 on the same (bf16-rounded) operands, then the
whole step (forward + L1 + backward + Adam) against an fp32 torch restatement of the reference loop body.
Tolerances: activations and activation gradients are bf16 (2^-8 relative per rounding), accumulation fp32; weight
gradients are compared by relative L2 error per parameter tensor (stated at each assert)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle_np as O
import pnnp_b200 as P
from pnnp_b200 import _lib, archs, train

pytestmark = pytest.mark.gpu
L = _lib


def _bf(t):
    return t.to(torch.bfloat16).float()


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def _sp():
    return L.stream_ptr(torch.device("cuda"))


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _ok():
    torch.cuda.synchronize()
    assert L.lib().pnnp_conv_pipeline_error() == 0 and L.lib().pnnp_wgrad_nhwc_pipeline_error() == 0


def test_l1_loss_and_gradient_match_torch():
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = (torch.rand((2, 4, 32, 48), device="cuda", generator=g) * 1.4 - 0.2).requires_grad_(True)
    hr = torch.rand((2, 4, 32, 48), device="cuda", generator=g)
    loss = F.l1_loss(pred.clamp(0, 1), hr)
    loss.backward()
    gp = torch.empty_like(hr)
    s = torch.zeros(1, dtype=torch.float64, device="cuda")
    L.check(L.lib().pnnp_l1_loss(pred.data_ptr(), hr.data_ptr(), gp.data_ptr(), pred.numel(), s.data_ptr(), _sp()), "l1")
    assert abs(s.item() / pred.numel() - loss.item()) < 1e-6
    assert torch.equal(gp, pred.grad)                      # +-1/N or 0: exact
    assert abs(O.l1_loss(pred.detach().cpu().numpy(), hr.cpu().numpy()) - s.item() / pred.numel()) < 1e-6


def test_adam_matches_torch_optim():
    g = torch.Generator(device="cuda").manual_seed(1)
    p0 = torch.randn(10007, device="cuda", generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for t in range(1, 4):
        gr = torch.randn(10007, device="cuda", generator=g)
        ref.grad = gr.clone()
        opt.step()
        L.check(L.lib().pnnp_adam_step(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-4, 0.9, 0.999,
                                       1e-8, t, 1.0, _sp()), "adam")
        assert (p - ref.detach()).abs().max().item() < 2e-7      # same formula, fp32; ulp-level differences only


@pytest.mark.parametrize("act", [L.ACT_LEAKY, L.ACT_RELU, L.ACT_NONE])
def test_act_backward_and_bias_gradient(act):
    g = torch.Generator(device="cuda").manual_seed(2)
    n, h, w, c = 2, 24, 40, 64
    out = _bf(torch.randn((n, h, w, c), device="cuda", generator=g))
    go = _bf(torch.randn((n, h, w, c), device="cuda", generator=g))
    gk = go.to(torch.bfloat16).clone()
    db = torch.zeros(c, device="cuda")
    ob = out.to(torch.bfloat16)
    L.check(L.lib().pnnp_act_bwd_bias(gk.data_ptr(), ob.data_ptr() if act else None, db.data_ptr(),
                                      n * h * w, c, act, _sp()), "act_bwd")
    slope = {L.ACT_LEAKY: torch.where(out > 0, 1.0, 0.2), L.ACT_RELU: (out > 0).float(), L.ACT_NONE: torch.ones_like(out)}[act]
    want = _bf(go * slope)
    assert torch.equal(gk.float(), want)
    assert (db - want.sum((0, 1, 2))).abs().max().item() < 1e-2


def test_maxpool_backward_with_skip():
    g = torch.Generator(device="cuda").manual_seed(3)
    n, h, w, c = 2, 16, 32, 32
    x = _bf(torch.randn((n, c, h, w), device="cuda", generator=g)).requires_grad_(True)
    gp = _bf(torch.randn((n, c, h // 2, w // 2), device="cuda", generator=g))
    gs = _bf(torch.randn((n, c, h, w), device="cuda", generator=g))
    F.max_pool2d(x, 2).backward(gp)
    gc = torch.empty((n, h, w, c), dtype=torch.bfloat16, device="cuda")
    gpn, xn, gsn = _nhwc(gp), _nhwc(x.detach()), _nhwc(gs)
    L.check(L.lib().pnnp_maxpool_bwd(gpn.data_ptr(), xn.data_ptr(), gsn.data_ptr(), gc.data_ptr(),
                                     n, h, w, c, L.ACT_NONE, _sp()), "pool_bwd")
    assert torch.equal(_nchw(gc), _bf(x.grad + gs))
    # fused LeakyReLU'(x): x is the activated output whose pooling is undone
    L.check(L.lib().pnnp_maxpool_bwd(gpn.data_ptr(), xn.data_ptr(), gsn.data_ptr(), gc.data_ptr(),
                                     n, h, w, c, L.ACT_LEAKY, _sp()), "pool_bwd")
    assert torch.equal(_nchw(gc), _bf(_bf(x.grad + gs) * torch.where(x.detach() > 0, 1.0, 0.2)))


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 32, 32), (8, 64, 64, 256), (3, 38, 52, 64)])
def test_maxpool_backward_with_bias_sums(n, h, w, c):
    """pnnp_maxpool_bwd_bias: gradient bit-identical to pnnp_maxpool_bwd, bias sums equal to the separate pass (pnnp_act_bwd_bias, act none)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    x = _nhwc(_bf(torch.randn((n, c, h, w), device="cuda", generator=g)))
    gp = _nhwc(_bf(torch.randn((n, c, h // 2, w // 2), device="cuda", generator=g)))
    gs = _nhwc(_bf(torch.randn((n, c, h, w), device="cuda", generator=g)))
    ref, gc = torch.empty_like(x), torch.empty_like(x)
    L.check(L.lib().pnnp_maxpool_bwd(gp.data_ptr(), x.data_ptr(), gs.data_ptr(), ref.data_ptr(), n, h, w, c, L.ACT_LEAKY, _sp()), "pool_bwd")
    want = torch.full((c,), 0.25, device="cuda")
    L.check(L.lib().pnnp_act_bwd_bias(ref.data_ptr(), None, want.data_ptr(), n * h * w, c, L.ACT_NONE, _sp()), "bias sums")
    db = torch.full((c,), 0.25, device="cuda")
    L.check(L.lib().pnnp_maxpool_bwd_bias(gp.data_ptr(), x.data_ptr(), gs.data_ptr(), gc.data_ptr(), db.data_ptr(),
                                          n, h, w, c, L.ACT_LEAKY, _sp()), "pool_bwd_bias")
    torch.cuda.synchronize()
    assert torch.equal(gc, ref)
    exact = ref.double().reshape(-1, c).sum(0) + 0.25
    tol = 1e-5 * ref.double().abs().reshape(-1, c).sum(0) + 1e-4
    assert ((db.double() - exact).abs() <= tol).all() and ((want.double() - exact).abs() <= tol).all()


@pytest.mark.parametrize("ci,co,h,w,n", [(16, 32, 16, 32, 1), (32, 32, 24, 40, 2), (32, 64, 20, 36, 1), (64, 64, 16, 48, 2),
                                         (64, 128, 16, 16, 2), (128, 64, 16, 32, 1), (128, 128, 24, 16, 1), (256, 256, 8, 16, 1),
                                         (512, 256, 8, 8, 1), (256, 512, 4, 6, 2)])
def test_wgrad_nhwc_3x3_matches_autograd(ci, co, h, w, n):
    """Weight gradient straight from NHWC operands (MN-major tcgen05 GEMM, csrc/wgrad_nhwc_tc.cu) incl. ragged pixel tiles."""
    g = torch.Generator(device="cuda").manual_seed(3 * ci + co)
    x = _bf(torch.randn((n, ci, h, w), device="cuda", generator=g))
    go = _bf(torch.randn((n, co, h, w), device="cuda", generator=g))
    wt = torch.zeros((co, ci, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(x, wt, padding=1).backward(go)
    gon, xn = _nhwc(go), _nhwc(x)
    dw = torch.zeros((9, ci, co), device="cuda")
    L.check(L.lib().pnnp_wgrad_nhwc(0, gon.data_ptr(), co, co, xn.data_ptr(), ci, ci, n, h, w, dw.data_ptr(), 0, ci, co, _sp()), "wgrad_nhwc")
    _ok()
    got = dw.permute(2, 1, 0).reshape(co, ci, 3, 3)
    assert _rel(got, wt.grad) < 1e-4, _rel(got, wt.grad)          # bf16 operands are exact in both; fp32 summation order only


def test_wgrad_nhwc_two_sources_accumulate_into_one_gradient():
    """torch.cat([up, skip], 1) -> conv: each source adds its own block of input-channel rows (ci_off)."""
    g = torch.Generator(device="cuda").manual_seed(9)
    n, h, w, c0, c1, co = 2, 16, 32, 32, 32, 32
    x0, x1 = (_bf(torch.randn((n, c, h, w), device="cuda", generator=g)) for c in (c0, c1))
    go = _bf(torch.randn((n, co, h, w), device="cuda", generator=g))
    wt = torch.zeros((co, c0 + c1, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(torch.cat([x0, x1], 1), wt, padding=1).backward(go)
    gon, x0n, x1n = _nhwc(go), _nhwc(x0), _nhwc(x1)
    dw = torch.zeros((9, c0 + c1, co), device="cuda")
    L.check(L.lib().pnnp_wgrad_nhwc(0, gon.data_ptr(), co, co, x0n.data_ptr(), c0, c0, n, h, w, dw.data_ptr(), 0, c0 + c1, co, _sp()), "wgrad")
    L.check(L.lib().pnnp_wgrad_nhwc(0, gon.data_ptr(), co, co, x1n.data_ptr(), c1, c1, n, h, w, dw.data_ptr(), c0, c0 + c1, co, _sp()), "wgrad")
    _ok()
    assert _rel(dw.permute(2, 1, 0).reshape(co, c0 + c1, 3, 3), wt.grad) < 1e-4


@pytest.mark.parametrize("ci,co,h,w", [(64, 32, 16, 32), (128, 64, 8, 24), (512, 256, 8, 8)])
def test_wgrad_nhwc_conv_transpose(ci, co, h, w):
    g = torch.Generator(device="cuda").manual_seed(ci + 1)
    n = 2
    x = _bf(torch.randn((n, ci, h, w), device="cuda", generator=g))
    wt = torch.zeros((ci, co, 2, 2), device="cuda", requires_grad=True)
    go = _bf(torch.randn((n, co, 2 * h, 2 * w), device="cuda", generator=g))
    F.conv_transpose2d(x, wt, stride=2).backward(go)
    gon, xn = _nhwc(go), _nhwc(x)
    dw = torch.zeros((4, ci, co), device="cuda")
    L.check(L.lib().pnnp_wgrad_nhwc(1, gon.data_ptr(), co, co, xn.data_ptr(), ci, ci, n, h, w, dw.data_ptr(), 0, ci, co, _sp()), "wgrad_nhwc")
    _ok()
    assert _rel(dw.permute(1, 2, 0).reshape(ci, co, 2, 2), wt.grad) < 1e-4


@pytest.mark.parametrize("ci,co,h,w", [(64, 32, 16, 32), (512, 256, 8, 8)])
def test_conv_transpose_backward_pieces(ci, co, h, w):
    """ConvTranspose2d(2, s2): data gradient = the 2x2 stride-2 mode of the conv kernel (wgrad: test_wgrad_nhwc_conv_transpose)."""
    g = torch.Generator(device="cuda").manual_seed(ci)
    n = 2
    x = _bf(torch.randn((n, ci, h, w), device="cuda", generator=g)).requires_grad_(True)
    wt = _bf(torch.randn((ci, co, 2, 2), device="cuda", generator=g) / ci ** 0.5).requires_grad_(True)
    go = _bf(torch.randn((n, co, 2 * h, 2 * w), device="cuda", generator=g))
    F.conv_transpose2d(x, wt, stride=2).backward(go)
    gx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device="cuda")
    archs._conv(L.CONV2S2, _nhwc(go), train._pack_conv_weight(wt.detach()), None, gx, ci, L.ACT_NONE)
    _ok()
    assert _rel(_nchw(gx), x.grad) < 6e-3                          # bf16 output rounding


def test_head_backward():
    g = torch.Generator(device="cuda").manual_seed(5)
    n, h, w, cin, co = 2, 16, 32, 32, 4
    pre = _bf(torch.randn((n, cin, h, w), device="cuda", generator=g))
    act = _bf(F.leaky_relu(pre, 0.2)).requires_grad_(True)
    wt = (torch.randn((co, cin, 1, 1), device="cuda", generator=g) * 0.1).requires_grad_(True)
    b = torch.zeros(co, device="cuda", requires_grad=True)
    gp = torch.randn((n, co, h, w), device="cuda", generator=g)
    F.conv2d(act, wt, b).backward(gp)
    gact = torch.empty((n, h, w, cin), dtype=torch.bfloat16, device="cuda")
    dW, db, dbp = torch.zeros((co, cin), device="cuda"), torch.zeros(co, device="cuda"), torch.zeros(cin, device="cuda")
    an, w2 = _nhwc(act.detach()), wt.detach().reshape(co, cin).contiguous()
    L.check(L.lib().pnnp_head_bwd(gp.data_ptr(), an.data_ptr(), w2.data_ptr(),
                                  gact.data_ptr(), dW.data_ptr(), db.data_ptr(), dbp.data_ptr(), n, h, w, cin, co, L.ACT_LEAKY, _sp()),
            "head")
    torch.cuda.synchronize()
    want = act.grad * torch.where(act.detach() > 0, 1.0, 0.2)
    assert _rel(_nchw(gact), want) < 6e-3
    assert _rel(dW, wt.grad.reshape(co, cin)) < 1e-4 and _rel(db, b.grad) < 1e-4
    assert _rel(dbp, want.sum((0, 2, 3))) < 1e-2


def _reference_step(sd, lr_in, hr, steps=1, lr=1e-4):
    """fp32 torch restatement of trainer_SID.py:93-101 on the oracle's functional UNet."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    losses, grads = [], None
    for _ in range(steps):
        opt.zero_grad()
        pred = O.unet_forward(lr_in, params)
        loss = F.l1_loss(pred.clamp(0, 1), hr)
        loss.backward()
        if grads is None:
            grads = {k: v.grad.clone() for k, v in params.items()}
        opt.step()
        losses.append(loss.item())
    return losses, grads, {k: v.detach() for k, v in params.items()}


def _make(n=2, h=64, w=96, seed=11, std=None):
    torch.manual_seed(seed)
    net = P.UNetSeeInDark({"in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": False}).cuda()
    archs.initialize_weights(net)
    if std is not None:            # He-like scaling so every layer carries signal and gradient of O(1)
        for m in net.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                fan = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
                m.weight.data.normal_(0, std / fan ** 0.5)
    g = torch.Generator(device="cuda").manual_seed(seed)
    hr = torch.rand((n, 4, h, w), device="cuda", generator=g) ** 2
    lr_in = (hr + 0.05 * torch.randn((n, 4, h, w), device="cuda", generator=g))
    return net, lr_in, hr


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()          # F.cosine_similarity clamps norms at 1e-8: useless for tiny gradients
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def _ste(t):
    """bf16 rounding with a straight-through gradient."""
    return t + (t.to(torch.bfloat16).float() - t).detach()


def _unet_forward_bf16_storage(x, sd):
    """archs/Unet.py:54-99 in fp32 arithmetic with every stored activation and every weight rounded to bf16 — the function
    the tcgen05 path computes (bf16 operands, fp32 accumulation), so max-pool routing and LeakyReLU masks are decided
    on the same values; autograd through it is the reference for the hand-written backward."""
    act = lambda t: _ste(F.leaky_relu(t, 0.2))
    cv = lambda t, n: F.conv2d(t, _ste(sd[n + ".weight"]), sd[n + ".bias"], padding=sd[n + ".weight"].shape[-1] // 2)
    up = lambda t, n: _ste(F.conv_transpose2d(t, _ste(sd[n + ".weight"]), sd[n + ".bias"], stride=2))
    c, cur = {}, _ste(x)
    for i in range(1, 6):
        c[i] = act(cv(act(cv(cur, f"conv{i}_1")), f"conv{i}_2"))
        cur = F.max_pool2d(c[i], 2) if i < 5 else c[i]
    for i in range(6, 10):
        cur = act(cv(act(cv(torch.cat([up(cur, f"upv{i}"), c[10 - i]], 1), f"conv{i}_1")), f"conv{i}_2"))
    return cv(cur, "conv10_1")


@pytest.mark.parametrize("std", [None, 1.4])
def test_training_step_gradients_match_fp32_autograd(std):
    """Backward pass in isolation: both references are driven by the SAME d loss / d pred (the L1 gradient is a sign
    function, so feeding each side its own would compare sign flips of a bf16-rounded prediction, not the backward)."""
    net, lr_in, hr = _make(std=std)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ts = train.UNetTrainStep(net)
    pred, saved = ts.forward(lr_in)
    gp = torch.empty_like(pred)
    L.check(L.lib().pnnp_l1_loss(pred.data_ptr(), hr.data_ptr(), gp.data_ptr(), pred.numel(), ts.loss_sum.data_ptr(), _sp()), "l1")
    ts.backward(gp, saved)
    _ok()
    # (1) plain fp32 network (the reference's arithmetic)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    pred_ref = O.unet_forward(lr_in, params)
    ref_loss = F.l1_loss(pred_ref.clamp(0, 1), hr).item()
    pred_ref.backward(gp)
    assert abs(ts.loss_sum.item() / pred.numel() - ref_loss) < 2e-3 * max(1.0, ref_loss)
    assert (pred - pred_ref).abs().max().item() < 2e-2 * max(1.0, pred_ref.abs().max().item())
    st32 = {name: (_rel(ts._grad_view(name), params[name].grad), _cos(ts._grad_view(name), params[name].grad)) for name in sd}
    # (2) fp32 autograd through the bf16-storage network (same routing decisions as the kernels)
    p16 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    _unet_forward_bf16_storage(lr_in, p16).backward(gp)
    st16 = {name: (_rel(ts._grad_view(name), p16[name].grad), _cos(ts._grad_view(name), p16[name].grad)) for name in sd}
    for tag, st in (("fp32", st32), ("bf16-storage", st16)):
        print(tag, "worst rel", max(st.items(), key=lambda kv: kv[1][0]), "worst cos", min(st.items(), key=lambda kv: kv[1][1]))
    # vs plain fp32: max-pool ties / activation signs decided on bf16-rounded values re-route a few percent of the deep
    # gradients (measured 2.7 % with the reference init, 9.8 % with O(1) activations): relative L2 < 15 %, cosine > 0.99
    bad = {k: v for k, v in st32.items() if not (v[1] > 0.99 and v[0] < 0.15)}
    assert not bad, bad
    # vs the bf16-storage network: the two bf16 forwards still drift apart (different fp32 summation orders -> different
    # bf16 roundings, amplified layer by layer when activations are O(1)), which flips ~0.3 % of the LeakyReLU masks:
    # measured 1.7 % (reference init) / 7.6 % (O(1) activations).  The exact statement is the layer-local test below.
    bad = {k: v for k, v in st16.items() if not (v[1] > 0.995 and v[0] < 0.10)}
    assert not bad, bad


def test_backward_is_exact_layer_by_layer_on_a_real_step():
    """Chain rule, link by link: every layer's weight / bias / input gradients of a real training step are compared with
    torch autograd of THAT layer evaluated on the step's own saved activations and incoming gradient, so routing decisions
    (LeakyReLU masks, max-pool arg-max) are shared and only rounding remains: weight and bias gradients to 5e-3 relative
    (fp32 accumulation of exact bf16 products), activation gradients to 6e-3 relative after bf16 rounding."""
    net, lr_in, hr = _make(std=1.4)
    ts = train.UNetTrainStep(net)
    pred, s = ts.forward(lr_in)
    gp = torch.empty_like(pred)
    L.check(L.lib().pnnp_l1_loss(pred.data_ptr(), hr.data_ptr(), gp.data_ptr(), pred.numel(), ts.loss_sum.data_ptr(), _sp()), "l1")
    ts.backward(gp, s)
    _ok()
    B = ts.scr.bufs
    W = {k: v.detach().clone() for k, v in net.state_dict().items()}
    slope = lambda a: torch.where(_nchw(a) > 0, 1.0, 0.2)               # LeakyReLU'(.) from the stored activated output
    worst = {"w": 0.0, "b": 0.0, "x": 0.0}

    def layer(name, xs, gz, conv):
        """xs: NHWC bf16 inputs; gz: NHWC bf16 gradient w.r.t. the layer's pre-activation.  Returns input gradients (NCHW fp32)."""
        xin = [_nchw(x).requires_grad_(True) for x in xs]
        w = _bf(W[name + ".weight"]).requires_grad_(True)
        b = W[name + ".bias"].clone().requires_grad_(True)
        cin_real = w.shape[1] if conv else w.shape[0]
        xcat = torch.cat(xin, 1)[:, :cin_real]                          # conv1_1: 16 stored channels, 4 real
        out = F.conv2d(xcat, w, b, padding=w.shape[-1] // 2) if conv else F.conv_transpose2d(xcat, w, b, stride=2)
        out.backward(_nchw(gz))
        rw, rb = _rel(ts._grad_view(name + ".weight"), w.grad), _rel(ts._grad_view(name + ".bias"), b.grad)
        worst["w"], worst["b"] = max(worst["w"], rw), max(worst["b"], rb)
        assert rw < 5e-3 and rb < 5e-3, (name, rw, rb)
        return [x.grad for x in xin]

    def check_x(tag, ours, want):
        r = _rel(_nchw(ours), _bf(want))
        worst["x"] = max(worst["x"], r)
        assert r < 6e-3, (tag, r)

    # 1x1 head (its kernel also applies conv9_2's LeakyReLU')
    (dx,) = layer("conv10_1", [s["c9"]], _nhwc(gp), True)
    check_x("head", B["g_c9"], dx * slope(s["c9"]))
    gz = B["g_c9"]
    for i in range(9, 5, -1):                                               # decoder
        (dx,) = layer(f"conv{i}_2", [s[f"c{i}a"]], gz, True)
        check_x(f"conv{i}_2.dx", B[f"gx_conv{i}_2_0"], dx * slope(s[f"c{i}a"]))
        dup, dskip = layer(f"conv{i}_1", [s[f"u{i}"], s[f"c{10 - i}"]], B[f"gx_conv{i}_2_0"], True)
        check_x(f"conv{i}_1.dx_up", B[f"gx_conv{i}_1_0"], dup)
        check_x(f"conv{i}_1.dx_skip", B[f"gx_conv{i}_1_1"], dskip)
        src = s["c5"] if i == 6 else s[f"c{i - 1}"]
        (dx,) = layer(f"upv{i}", [src], B[f"gx_conv{i}_1_0"], False)
        check_x(f"upv{i}.dx", B[f"gx_upv{i}"], dx * slope(src))
        gz = B[f"gx_upv{i}"]
    for i in range(5, 0, -1):                                               # encoder
        (dx,) = layer(f"conv{i}_2", [s[f"c{i}a"]], gz, True)
        check_x(f"conv{i}_2.dx", B[f"gx_conv{i}_2_0"], dx * slope(s[f"c{i}a"]))
        (dx,) = layer(f"conv{i}_1", [s[f"in{i}_1"]], B[f"gx_conv{i}_2_0"], True)
        if i > 1:
            check_x(f"conv{i}_1.dx", B[f"gx_conv{i}_1_0"], dx)
            c = _nchw(s[f"c{i - 1}"]).requires_grad_(True)                  # max-pool routing + skip add + LeakyReLU'
            F.max_pool2d(c, 2).backward(_nchw(B[f"gx_conv{i}_1_0"]))
            check_x(f"pool{i - 1}", B[f"g_c{i - 1}"], _bf(c.grad + _nchw(B[f"gx_conv{11 - i}_1_1"])) * slope(s[f"c{i - 1}"]))
            gz = B[f"g_c{i - 1}"]
    print("worst relative errors:", worst)


def test_training_reduces_loss_like_the_reference_loop():
    """Ten Adam steps (lr 1e-3) on fixed crops: the loss curve follows the fp32 reference loop and the parameters move with it."""
    net, lr_in, hr = _make(n=2, h=64, w=64, seed=5)
    net.conv10_1.bias.data.fill_(0.05)       # all four outputs start inside the clamp (a negative bias means zero gradient)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ref_losses, _, ref_params = _reference_step(sd, lr_in, hr, steps=10, lr=1e-3)
    ts = train.UNetTrainStep(net, lr=1e-3)
    losses = [ts.step(lr_in, hr).item() for _ in range(10)]
    _ok()
    assert losses[-1] < 0.97 * losses[0] and ref_losses[-1] < 0.97 * ref_losses[0], (losses, ref_losses)
    assert np.allclose(losses, ref_losses, rtol=3e-2, atol=2e-3), (losses, ref_losses)
    moved = max((v - sd[k]).abs().max().item() for k, v in net.state_dict().items())
    assert 5e-3 < moved < 1.5e-2                                   # ~ steps * lr, as Adam's first steps do
    # sign-like early Adam steps: weights whose tiny gradient flips sign between bf16 and fp32 drift apart by 2 lr per step, so
    # the bound is on the bulk: 99 % of all parameters within 20 % of the distance travelled
    diff = torch.cat([(v - ref_params[k]).abs().flatten() for k, v in net.state_dict().items()])
    assert torch.quantile(diff[::7].float(), 0.99).item() < 0.2 * moved, torch.quantile(diff[::7].float(), 0.99).item()
    # the inference forward sees the updated weights (pack cache invalidated)
    with torch.no_grad():
        out = net.eval()(lr_in)
    assert (out - O.unet_forward(lr_in, {k: v for k, v in net.state_dict().items()})).abs().max().item() < 1e-2


def test_graph_replay_takes_the_same_steps_as_eager_launches():
    """The captured CUDA graph of a step (device-side learning rate / step count) == the same kernels launched one by one; the
    learning rate can change between replays."""
    outs = {}
    for use_graph in (False, True):
        net, lr_in, hr = _make(n=2, h=64, w=64, seed=5)
        net.conv10_1.bias.data.fill_(0.05)
        ts = train.UNetTrainStep(net, lr=1e-3)
        ts.use_graph = use_graph
        losses = []
        for i in range(6):
            if i == 3:
                ts.lr = 5e-4
            losses.append(ts.step(lr_in, hr).item())
        _ok()
        assert (ts._graphs[((2, 4, 64, 64), True)]["graph"] is not None) == use_graph
        assert abs(ts.adam_state[1].item() - 6.0) < 1e-6 and abs(ts.adam_state[0].item() - 5e-4) < 1e-9
        outs[use_graph] = (losses, ts.flat_p.clone())
    assert np.allclose(outs[False][0], outs[True][0], rtol=1e-4, atol=1e-6), outs
    # split-K / bias sums use fp32 atomics, so the two runs agree to summation order, not bit for bit: the bulk of the parameters
    # within 2e-4; a weight whose gradient is numerically zero may take sign-like Adam steps in either direction (<= steps * lr)
    diff = (outs[False][1] - outs[True][1]).abs()
    assert torch.quantile(diff[::5].float(), 0.99).item() < 2e-4 and diff.max().item() < 6.5e-3
