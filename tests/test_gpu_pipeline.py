"""pipeline.SynthDenoisePipeline — the whole hot path behind one host-buffer call (pinned uint16 RAW crops in, PSNR/SSIM partial sums
out; what bench.py times as e2e) — against the same stages called one by one, and against the oracle on a crop."""
import numpy as np
import pytest
import torch

import oracle_np as O
import pnnp_b200 as P
from pnnp_b200 import _lib
from pnnp_b200.metrics import eval_partial_sums, finish_metrics
from pnnp_b200.pipeline import SynthDenoisePipeline

pytestmark = pytest.mark.gpu
ARCH = dict(name="UNetSeeInDark", in_nc=4, out_nc=4, nf=32, nframes=1, use_dpsv=False, res=False, cascade=False, add=False, lock_wb=False)


def test_pipeline_equals_the_stages_called_one_by_one():
    torch.manual_seed(2)
    net = P.UNetSeeInDark(ARCH).cuda().eval()
    P.initialize_weights(net)
    n, H, W = 10, 128, 192                                            # 10 crops, chunks of 4: a ragged last chunk
    rs = np.random.RandomState(4)
    batches = [torch.from_numpy(rs.randint(512, 16384, size=(n, H, W)).astype(np.uint16).view(np.int16)).pin_memory() for _ in range(3)]
    np.random.seed(7)
    params = [P.sample_params("SonyA7S2") for _ in range(n)]
    pipe = SynthDenoisePipeline(net, n, H, W, 16383, 512, "pgrq", chunk=4)
    got = []
    for i, b in enumerate(batches):                                  # the next batch is announced one call ahead (prefetched H2D)
        nxt = batches[i + 1] if i + 1 < len(batches) else None
        sums = pipe.run(b, params=params, seed_offset=(11, i), crop_id0=100, next_host=nxt)
        torch.cuda.synchronize()
        got.append(sums.clone())
    table = P.ParamTable(params, "cuda")
    for i, b in enumerate(batches):
        with torch.no_grad():
            hr = P.raw2bayer(b.cuda(), wp=16383, bl=512, norm=True, clip=True)
            lr = P.synthesize_batch(hr, None, "pgrq", post_clip=(-float("inf"), 1.0), crop_id0=100, table=table, seed_offset=(11, i))
            want = eval_partial_sums(net(lr), hr, 1.0, False).cpu()
        # same kernels, same draws (keyed on global crop ids); the sums are float64 atomics, i.e. equal to summation order
        assert torch.allclose(got[i], want, rtol=1e-11, atol=1e-9), i
    # a batch that was NOT announced still goes through (the stale prefetch is overwritten in stream order)
    other = torch.from_numpy(rs.randint(512, 16384, size=(n, H, W)).astype(np.uint16).view(np.int16)).pin_memory()
    pipe.run(batches[0], params=params, seed_offset=(11, 0), crop_id0=100, next_host=batches[1])
    sums = pipe.run(other, params=params, seed_offset=(11, 5), crop_id0=100)
    torch.cuda.synchronize()
    with torch.no_grad():
        hr = P.raw2bayer(other.cuda(), wp=16383, bl=512, norm=True, clip=True)
        lr = P.synthesize_batch(hr, None, "pgrq", post_clip=(-float("inf"), 1.0), crop_id0=100, table=table, seed_offset=(11, 5))
        want = eval_partial_sums(net(lr), hr, 1.0, False).cpu()
    assert torch.allclose(sums, want, rtol=1e-11, atol=1e-9)


def test_pipeline_metrics_against_the_oracle_on_one_crop():
    """PSNR / SSIM of the denoised crop as the oracle computes them from the pipeline's own noisy crop (the draws are Philox's, so
    the noisy crop is taken from the device; everything after it is the oracle's: fp32 UNet, clamp, tensor2im, PSNR, SSIM)."""
    torch.manual_seed(3)
    net = P.UNetSeeInDark(ARCH).cuda().eval()
    P.initialize_weights(net)
    H, W = 128, 160
    raw = np.random.RandomState(1).randint(512, 9000, size=(1, H, W)).astype(np.uint16)
    np.random.seed(1)
    params = [P.sample_params("SonyA7S2")]
    pipe = SynthDenoisePipeline(net, 1, H, W, 16383, 512, "pgrq", chunk=1)
    sums = pipe.run(torch.from_numpy(raw.view(np.int16)).pin_memory(), params=params, seed_offset=(5, 0))
    torch.cuda.synchronize()
    got = finish_metrics(sums, 4, H // 2, W // 2)[0]
    hr = O.raw2bayer(raw[0], wp=16383, bl=512, norm=True, clip=True)
    lr = P.synthesize_batch(torch.from_numpy(hr).cuda()[None], params, "pgrq", post_clip=(-float("inf"), 1.0), seed_offset=(5, 0))
    with torch.no_grad():
        dn = O.unet_forward(lr.cpu(), {k: v.cpu() for k, v in net.state_dict().items()}).clamp(0, 1).numpy()
    a, b = O.tensor2im(dn), O.tensor2im(hr[None])
    assert got["PSNR"] == pytest.approx(O.psnr(b, a), abs=0.01)      # north_star: PSNR within 0.01 dB on the same inputs
    assert got["SSIM"] == pytest.approx(O.ssim(b, a), abs=1e-4)
