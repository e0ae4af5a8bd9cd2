"""Super-tile variant of the conv kernel (conv_tc.cu, template parameter SUP; PNNP_CONV_SUPER=1|2): two M = 128 tiles per pipeline
stage.  Every output pixel sees the same MMAs in the same order as in the default kernel, so the two must agree BIT FOR BIT —
outputs, fused max-pool, fused 1x1 head and the masked data-gradient epilogue alike.

The variant was written after the round's GPU budget was spent and has not run on a B200 yet: it is opt-in in the product
(environment variable, default off) and these tests are opt-in too (PNNP_TEST_EXPERIMENTAL=1), so the default `pytest -m gpu`
run exercises only measured code."""
import os

import pytest
import torch

import pnnp_b200 as P
from pnnp_b200 import _lib, archs

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PNNP_TEST_EXPERIMENTAL") != "1", reason="experimental kernel variant: opt-in")]


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
        monkeypatch.delenv("PNNP_CONV_SUPER", raising=False)
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
