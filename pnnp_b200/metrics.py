"""Eval metrics behind the reference's names (utils/visualization.py:9-31,
data_process/__init__.py:144-175), computed on the device by csrc/eval_metrics.cu.

`eval_frame_metrics` replaces the sequence  clamp -> IlluminanceCorrect -> tensor2im x2 ->
quality_assess  of trainer_SID.py:231-248: no 2 x 48.5 MB device->host copy, no host SSIM."""
import math

import torch

from . import _lib


def eval_partial_sums(dn, hr, scale=1.0, brightness_correct=False):
    """dn, hr: CUDA fp32 (n,c,h,w).  Returns a CUDA float64 tensor (n, 3+c) of partial sums
    [num, den, sse, ssim_sum_c...] (see include/pnnp_b200.h)."""
    _lib.require_cuda(dn, "dn")
    _lib.require_cuda(hr, "hr")
    if dn.shape != hr.shape or dn.dtype != torch.float32 or hr.dtype != torch.float32:
        raise RuntimeError("pnnp_b200: dn and hr must be float32 tensors of equal shape")
    n, c, h, w = dn.shape
    sums = torch.empty((n, 3 + c), dtype=torch.float64, device=dn.device)
    with torch.cuda.device(dn.device):
        _lib.check(_lib.lib().pnnp_eval_epilogue(dn.data_ptr(), hr.data_ptr(), n, c, h, w, float(scale),
                                                 int(bool(brightness_correct)), sums.data_ptr(),
                                                 _lib.stream_ptr(dn.device)), "eval_epilogue")
    return sums


def finish_metrics(sums, c, h, w):
    """Partial sums (host or device tensor, (n, 3+c)) -> list of {'PSNR', 'SSIM'} per frame."""
    s = sums.detach().cpu().tolist()
    out = []
    for row in s:
        mse = row[2] / (c * h * w)
        psnr = float("inf") if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)
        valid = (h - 6) * (w - 6)
        ssim = sum(row[3:3 + c]) / (c * valid)
        out.append({"PSNR": psnr, "SSIM": ssim})
    return out


def eval_frame_metrics(dn, hr, scale=1.0, brightness_correct=False):
    n, c, h, w = dn.shape
    return finish_metrics(eval_partial_sums(dn, hr, scale, brightness_correct), c, h, w)


class IlluminanceCorrect(torch.nn.Module):
    """data_process/__init__.py:144-175 as a module (returns the corrected image, like the reference).
    The gain comes from the device partial sums; only one scalar multiply runs in torch."""

    def forward(self, predict, source):
        outs = []
        for i in range(predict.shape[0]):
            src = source[i:i + 1] if source.shape[0] != 1 else source
            p = torch.clamp(predict[i:i + 1], 0, 1).contiguous()
            sums = eval_partial_sums(p, src.contiguous().float(), 1.0, True)
            gain = (sums[0, 0].float() / sums[0, 1].float())
            outs.append(gain * p)
        return torch.cat(outs, 0)


def quality_assess(X, Y, data_range=255):
    """utils/visualization.py:26-31 signature; X estimate, Y target as CUDA NCHW [0,1] tensors.
    (The reference passes HWC [0,255] NumPy arrays produced by tensor2im; here tensor2im is fused.)"""
    if data_range != 255:
        raise NotImplementedError("only data_range=255 (the reference's call) is implemented")
    return eval_frame_metrics(X, Y)[0]
