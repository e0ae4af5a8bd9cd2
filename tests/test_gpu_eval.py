"""E1/E2 on the device vs the oracle restatement (and the reference's IlluminanceCorrect golden).
skimage is not installed in the build image, so PSNR/SSIM parity is pinned only to the oracle's
restatement of skimage's documented defaults ("parity unpinned" for E2)."""
import numpy as np
import pytest
import torch

import oracle_np as O
from pnnp_b200 import metrics

pytestmark = pytest.mark.gpu


def _oracle_metrics(dn, hr, scale, correct):
    d = torch.clamp(torch.from_numpy(dn) * scale, 0, 1)
    t = torch.from_numpy(hr)
    if correct:
        d = O.illuminance_correct(d, t)
    a, b = O.tensor2im(d.numpy()), O.tensor2im(hr)
    return O.psnr(b, a), O.ssim(b, a)


@pytest.mark.parametrize("shape", [(1, 4, 40, 72), (1, 4, 67, 131), (2, 4, 96, 64)])
@pytest.mark.parametrize("correct", [False, True])
def test_psnr_ssim_vs_oracle(shape, correct):
    rs = np.random.RandomState(shape[2])
    hr = rs.rand(*shape).astype(np.float32)
    hr[:, 0, :2, :7] = 1.0                                       # saturated pixels are excluded from the gain
    dn = (hr * 0.9 + 0.05 * rs.randn(*shape)).astype(np.float32)
    res = metrics.eval_frame_metrics(torch.from_numpy(dn).cuda(), torch.from_numpy(hr).cuda(), 1.0, correct)
    for i in range(shape[0]):
        p, s = _oracle_metrics(dn[i:i + 1], hr[i:i + 1], 1.0, correct)
        assert res[i]["PSNR"] == pytest.approx(p, abs=2e-4), (res[i], p)
        assert res[i]["SSIM"] == pytest.approx(s, abs=1e-6), (res[i], s)


def test_ratio_scale_and_identical_images():
    rs = np.random.RandomState(1)
    hr = rs.rand(1, 4, 48, 48).astype(np.float32)
    dn = hr / 100.0
    r = metrics.eval_frame_metrics(torch.from_numpy(dn).cuda(), torch.from_numpy(hr).cuda(), 100.0, False)[0]
    p, s = _oracle_metrics(dn, hr, 100.0, False)
    assert r["PSNR"] == pytest.approx(p, abs=1e-3) and r["SSIM"] == pytest.approx(s, abs=1e-6)
    same = metrics.eval_frame_metrics(torch.from_numpy(hr).cuda(), torch.from_numpy(hr).cuda())[0]
    assert same["PSNR"] == float("inf") and same["SSIM"] == pytest.approx(1.0, abs=1e-12)


def test_illuminance_correct_vs_reference_golden(golden):
    g = golden("eval")
    out = metrics.IlluminanceCorrect()(torch.from_numpy(g["pred"]).cuda(), torch.from_numpy(g["src"]).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), g["corrected"], rtol=2e-6, atol=1e-7)


def test_full_frame_vs_oracle():
    """E1 + E2 at the BASELINE frame shape (1 x 4 x 1424 x 2128, configs[0] / [4]) against the oracle — clamp, IlluminanceCorrect
    gain, tensor2im, PSNR and SSIM (trainer_SID.py:231-244) — not just a plausible range."""
    g = torch.Generator(device="cuda").manual_seed(0)
    hr = torch.rand((1, 4, 1424, 2128), device="cuda", generator=g)
    hr[:, 1, :3, :11] = 1.0                                      # saturated samples: excluded from the gain's dot products
    dn = (0.93 * hr + 0.02 * torch.randn(hr.shape, device="cuda", generator=g)).contiguous()
    for correct in (True, False):
        r = metrics.eval_frame_metrics(dn, hr, 1.0, correct)[0]
        p, s = _oracle_metrics(dn.cpu().numpy(), hr.cpu().numpy(), 1.0, correct)
        print(f"full frame correct={correct}: PSNR {r['PSNR']:.6f} (oracle {p:.6f}), SSIM {r['SSIM']:.8f} (oracle {s:.8f})")
        assert r["PSNR"] == pytest.approx(p, abs=2e-4) and r["SSIM"] == pytest.approx(s, abs=1e-6)
