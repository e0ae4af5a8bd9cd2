"""Bayer pack / unpack behind the reference's signatures (utils/isp_ops.py:84-112), executed by
the sm_100a kernels in csrc/pack.cu.

NumPy in → NumPy out (host↔device copies included, drop-in for the reference's call sites);
CUDA tensor in → CUDA tensor out (no copies; what the fused pipeline uses)."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _device():
    return _lib.cuda_device()


def raw2bayer(raw, wp=1023, bl=64, norm=True, clip=False, bias=np.array([0, 0, 0, 0])):
    """utils/isp_ops.py:84-96.  raw: H×W (or n×H×W) uint16 / float → float32 (n×)4×H/2×W/2,
    planes R(0,0) G1(0,1) B(1,1) G2(1,0); (x-(bias+bl))/(wp-(bias+bl)) in float64, rounded once."""
    host = not isinstance(raw, torch.Tensor)
    if host:
        a = np.asarray(raw)
        if a.dtype == np.uint16:
            t = torch.from_numpy(a.astype(np.int16, copy=False).view(np.int16)).to(_device())  # bit-preserving
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_device())
        is_u16 = a.dtype == np.uint16
    else:
        t = raw
        is_u16 = t.dtype in (torch.int16, torch.uint16)
        if not is_u16:
            t = t.float()
        t = t.contiguous()
    _lib.require_cuda(t, "raw")
    batched = t.dim() == 3
    n = t.shape[0] if batched else 1
    H, W = t.shape[-2], t.shape[-1]
    out = torch.empty((n, 4, H // 2, W // 2), dtype=torch.float32, device=t.device)
    black = (C.c_double * 4)(*[float(b) + float(bl) for b in np.asarray(bias).reshape(-1)[:4]])
    fn = _lib.lib().pnnp_pack_norm_u16 if is_u16 else _lib.lib().pnnp_pack_norm_f32
    with torch.cuda.device(t.device):
        _lib.check(fn(t.data_ptr(), out.data_ptr(), n, H, W, float(wp), black, int(bool(norm)), int(bool(clip)),
                      _lib.stream_ptr(t.device)), "raw2bayer")
    if not batched:
        out = out[0]
    return out.cpu().numpy() if host else out


def bayer2raw(packed_raw, wp=16383, bl=512, device_out=False):
    """utils/isp_ops.py:98-112.  (1×)4×h×w float → 2h×2w uint16 (clip [0,1], x*(wp-bl)+bl in
    float32, truncating cast).  Returns a NumPy array like the reference unless device_out."""
    if isinstance(packed_raw, torch.Tensor):
        t = packed_raw.detach()
        if t.dim() == 4:
            t = t[0]
        t = t.float().to(_device()) if not t.is_cuda else t.float()
    else:
        a = np.asarray(packed_raw)
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_device())
    t = t.contiguous()
    _, h, w = t.shape
    raw = torch.empty((2 * h, 2 * w), dtype=torch.int16, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().pnnp_unpack_quant(t.data_ptr(), raw.data_ptr(), 1, h, w, float(wp), float(bl),
                                                _lib.stream_ptr(t.device)), "bayer2raw")
    if device_out:
        return raw.view(torch.uint16)
    return raw.cpu().numpy().view(np.uint16)


def bayer2rggb(bayer):
    """utils/isp_ops.py:57-59 (host view op; true RGGB HWC order — differs from raw2bayer)."""
    H, W = bayer.shape
    return bayer.reshape(H // 2, 2, W // 2, 2).transpose(0, 2, 1, 3).reshape(H // 2, W // 2, 4)


def rggb2bayer(rggb):
    """utils/isp_ops.py:61-63."""
    H, W, _ = rggb.shape
    return rggb.reshape(H, W, 2, 2).transpose(0, 2, 1, 3).reshape(H * 2, W * 2)
