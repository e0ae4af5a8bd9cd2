"""Real-data side of the noise path, on the device:

* `darkshading_raw2bayer` — dark-shading subtraction fused into the Bayer pack (data_process/real_datasets.py:360-372:
  `lr_raw - get_darkshading(iso) [+ mean] [+ randn * biassig]` followed by raw2bayer(norm=True, clip=False));
* `HighBitRecovery` — the reference's class (data_process/process.py:675-751): the LUT is built on the host exactly as
  there (scipy.stats cdf per integer DN level), `map` runs on the device (csrc/hbr.cu) with Philox draws, or with
  caller-supplied uniforms to replay the reference's own draws.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .noise_params import sample_params_max
from .rng import default_generator


def darkshading_raw2bayer(lr_raw, darkshading, wp=16383, bl=512, add_mean=False, bias_draw=None, clip=False,
                          bias=np.array([0, 0, 0, 0])):
    """lr_raw: H x W (or n x H x W) uint16 sensor frame (NumPy array or CUDA int16/uint16 tensor); darkshading: H x W float32 or
    float64 map (`ds_k * iso + ds_b + BLE`).  Returns the packed float32 CUDA tensor raw2bayer would give for
    `lr_raw - darkshading [+ darkshading.mean()] [+ bias_draw]`, evaluated in the map's precision as NumPy does."""
    dev = _lib.cuda_device()                                    # raises without a CUDA device: there is no CPU fallback
    if isinstance(lr_raw, torch.Tensor):
        t = lr_raw
        if t.dtype not in (torch.int16, torch.uint16):
            raise RuntimeError("darkshading_raw2bayer: the sensor frame must be uint16")
    else:
        a = np.ascontiguousarray(lr_raw)
        if a.dtype != np.uint16:
            raise RuntimeError("darkshading_raw2bayer: the sensor frame must be uint16")
        t = torch.from_numpy(a.view(np.int16)).to(dev)
    t = t.contiguous()
    _lib.require_cuda(t, "lr_raw")
    if isinstance(darkshading, torch.Tensor):
        d = darkshading
    else:
        d = torch.from_numpy(np.ascontiguousarray(darkshading))
    if d.dtype not in (torch.float32, torch.float64):
        d = d.double()
    d = d.to(t.device).contiguous()
    batched = t.dim() == 3
    n = t.shape[0] if batched else 1
    H, W = t.shape[-2], t.shape[-1]
    if tuple(d.shape) != (H, W):
        raise RuntimeError(f"darkshading_raw2bayer: map {tuple(d.shape)} vs frame {(H, W)}")
    mean = 0.0
    if add_mean:           # `darkshading.mean()` keeps the map's dtype (np.float32 / np.float64 scalar)
        mean = float(d.mean(dtype=d.dtype))
    out = torch.empty((n, 4, H // 2, W // 2), dtype=torch.float32, device=t.device)
    black = (C.c_double * 4)(*[float(b) + float(bl) for b in np.asarray(bias).reshape(-1)[:4]])
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().pnnp_pack_norm_dark_u16(t.data_ptr(), d.data_ptr(), int(d.dtype == torch.float64), out.data_ptr(), n, H, W,
                                                      float(wp), black, 1, int(bool(clip)), mean, int(bool(add_mean)),
                                                      0.0 if bias_draw is None else float(bias_draw), int(bias_draw is not None),
                                                      _lib.stream_ptr(t.device)), "darkshading_raw2bayer")
    return out if batched else out[0]


class HighBitRecovery:
    """data_process/process.py:675-751 — same constructor, `get_lut`, `HB2LB_LUT` (host, scipy) and `map` (device)."""

    def __init__(self, camera_type='IMX686', noise_code='prq', param=None, perturb=True, factor=6, float=True):
        self.camera_type = camera_type
        self.noise_code = noise_code
        self.param = param
        self.perturb = perturb
        self.factor = factor
        self.float = float
        self.lut = {}

    def get_lut(self, iso_list, blc_mean=None):
        for iso in iso_list:
            bias = 0 if blc_mean is None else np.mean(blc_mean[iso])
            if self.perturb:
                bias += np.random.randn() * 0.1
            self.lut[iso] = self.HB2LB_LUT(iso, bias)

    def HB2LB_LUT(self, iso, bias=0, param=None):
        from scipy import stats
        lut_info = {}
        p = sample_params_max(self.camera_type, iso=iso) if param is None else param
        lut_info['param'] = p
        if 'g' in self.noise_code.lower():
            dist = stats.tukeylambda(p['lam'], loc=bias, scale=p['sigTL'])
            sigma = p['sigTL']
        else:
            dist = stats.norm(loc=bias, scale=p['sigGs'])
            sigma = p['sigGs']
        lut_info['dist'] = dist
        low = max(int(-sigma * self.factor + bias), -p['bl'] + 1)
        high = int(sigma * self.factor + bias)
        for x in range(low, high):
            lut_info[x] = {'cdf': dist.cdf(x - 0.5), 'range': dist.cdf(x + 0.5) - dist.cdf(x - 0.5)}
        lut_info.update(low=low, high=high, bias=bias, sigma=sigma)
        # device copy of the table (not part of the reference's dict)
        n = max(high - low, 1)
        tab = np.zeros((2, n), dtype=np.float64)
        for x in range(low, high):
            tab[0, x - low], tab[1, x - low] = lut_info[x]['cdf'], lut_info[x]['range']
        lut_info['_table'] = tab
        return lut_info

    def map(self, data, iso=6400, norm=True, rand=None, generator=None, index0=0, return_rand=False):
        """data: CUDA float32 tensor (any shape), normalised ([.., 1]) or in DN.  rand: optional float64 tensor of U(0,1)
        draws of data's shape (the reference's `np.random.uniform(0, 1, size=data.shape)`); otherwise Philox."""
        _lib.require_cuda(data, "data")
        lut = self.lut[iso]
        p = lut['param']
        x = data.float().contiguous()
        span = float(p['wp'] - p['bl'])
        scale_in = bool((x.max() <= 1).item())                       # `if np.max(data) <= 1`
        tab = torch.from_numpy(lut['_table']).to(x.device)
        out = torch.empty_like(x)
        tukey = 'g' in self.noise_code.lower()
        sc = float(p['sigTL'] if tukey else p['sigGs'])
        if rand is not None:
            rand = rand.to(device=x.device, dtype=torch.float64).contiguous()
            if rand.numel() != x.numel():
                raise RuntimeError("HighBitRecovery.map: rand must have data's shape")
            seed = offset = 0
        else:
            seed, offset = (default_generator if generator is None else generator).next()
        r_out = torch.empty(x.shape, dtype=torch.float64, device=x.device) if return_rand else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().pnnp_hbr_map(x.data_ptr(), out.data_ptr(), x.numel(), tab[0].data_ptr(), tab[1].data_ptr(),
                                               int(lut['low']), int(lut['high']), int(scale_in), int(bool(norm)) | (0 if self.float else 2), span, float(p['bl']),
                                               int(tukey), float(p['lam']), float(lut['bias']), sc, _lib.ptr(rand), seed, offset,
                                               int(index0), _lib.ptr(r_out), _lib.stream_ptr(x.device)), "hbr_map")
        return (out, r_out) if return_rand else out
