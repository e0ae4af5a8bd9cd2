"""Crop + augmentation of SynBase_Dataset (data_process/syn_datasets.py:69-107,162-173) on the device.

`init_random_crop_point` reproduces the reference's draw order on NumPy's global RandomState
(aug modes first, then per crop h_start, w_start), so the same seed gives the same crops;
`random_crop` runs the gather kernel (csrc/crop_aug.cu).  `eval_crop` / `eval_merge` are the overlapped tiling of a full frame
for tile-wise inference and its inverse (syn_datasets.py:109-159)."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def init_random_crop_point(h, w, patch_size, crop_per_image, mode='random_crop'):
    """syn_datasets.py:69-98 -> (h_start, w_start, aug) lists."""
    aug = np.random.randint(8, size=crop_per_image)
    hs, ws = [], []
    if mode == 'non-overlapped':
        nh, nw = h // patch_size, w // patch_size
        h0 = np.random.randint(0, h - nh * patch_size + 1)
        w0 = np.random.randint(0, w - nw * patch_size + 1)
        for i in range(nh):
            for j in range(nw):
                hs.append(h0 + i * patch_size)
                ws.append(w0 + j * patch_size)
    else:
        for _ in range(crop_per_image):
            hs.append(np.random.randint(0, h - patch_size + 1))
            ws.append(np.random.randint(0, w - patch_size + 1))
    return hs, ws, aug


def data_aug(data, mode=0):
    """syn_datasets.py:100-107 on a CUDA tensor (c,h,w) — via the same kernel with a full-frame crop."""
    c, h, w = data.shape
    if h != w:
        raise RuntimeError("data_aug: square crops only (rot90 of a non-square crop changes its shape)")
    return random_crop(data, [0], [0], [mode], h)[0]


def random_crop(img, h_start, w_start, aug, patch_size):
    """syn_datasets.py:162-173: img CUDA fp32 (c,h,w) -> crops (crop_per_image,c,patch,patch), crop_per_image = len(aug)
    (init_random_crop_point draws exactly that many aug modes).  As in the reference the first crop_per_image start points
    are used — a 'non-overlapped' grid may hold more — and fewer start points than crops raise IndexError."""
    _lib.require_cuda(img, "img")
    img = img.float().contiguous()
    c, h, w = img.shape
    n = k = len(aug)
    if len(h_start) < n or len(w_start) < n:
        raise IndexError(f"random_crop: {n} crops per image but only {min(len(h_start), len(w_start))} crop points")
    out = torch.empty((n, c, patch_size, patch_size), dtype=torch.float32, device=img.device)
    L = _lib.lib()
    with torch.cuda.device(img.device):
        for s in range(0, n, 64):
            m = min(64, n - s)
            arr = lambda v: (C.c_int * m)(*[int(x) for x in v[s:s + m]])
            modes = [int(aug[i]) if i < k else 0 for i in range(s, s + m)]
            _lib.check(L.pnnp_crop_aug(img.data_ptr(), out[s:].data_ptr(), c, h, w, patch_size, m, arr(h_start), arr(w_start),
                                       (C.c_int * m)(*modes), _lib.stream_ptr(img.device)), "crop_aug")
    return out


def wb_jitter(hr_crops, wb, gains):
    """syn_datasets.py:314-319 on CUDA crops (n,c,h,w), IN PLACE: `hr_crops *= rgb_gain`, then planes 0 and 2 times
    `wb[0] / red_gain` and `wb[2] / blue_gain`.  `gains` = random_gains().  The effective gains are formed on the host with the
    reference's own NumPy expression, so their dtype — float32 for a float32 / python-float white balance, float64 for an
    np.float64 one (NEP 50) — selects the float32 or the float64 product on the device exactly as NumPy would."""
    _lib.require_cuda(hr_crops, "hr_crops")
    if hr_crops.dtype != torch.float32:
        raise RuntimeError("wb_jitter: float32 crops")
    n, c, h, w = hr_crops.shape
    rgb_gain, red_gain, blue_gain = (g.numpy() if isinstance(g, torch.Tensor) else np.asarray(g) for g in gains)
    eff = {0: wb[0] / red_gain, 2: wb[2] / blue_gain}
    kind, gain = [0] * c, [1.0] * c
    for ch, g in eff.items():
        g = np.asarray(g)
        kind[ch] = 2 if g.dtype == np.float64 else 1
        gain[ch] = float(g.reshape(-1)[0])                     # float32 -> python float is exact
    with torch.cuda.device(hr_crops.device):
        _lib.check(_lib.lib().pnnp_wb_gains(hr_crops.data_ptr(), n, c, h, w, float(np.float32(rgb_gain.reshape(-1)[0])),
                                            (C.c_int * c)(*kind), (C.c_double * c)(*gain), _lib.stream_ptr(hr_crops.device)),
                   "wb_gains")
    return hr_crops


def tile_grid(h, w, patch_size, base=64):
    """(nh, nw) of SynBase_Dataset.eval_crop (syn_datasets.py:112-115)."""
    l = patch_size - base
    return h // l + 1, w // l + 1


def eval_crop(data, patch_size, base=64):
    """syn_datasets.py:109-133: data CUDA fp32 (1,c,h,w) or (c,h,w) -> (nh*nw, c, patch, patch) overlapped tiles of the
    reflect-padded frame."""
    _lib.require_cuda(data, "data")
    x = data.float().contiguous()
    if x.dim() == 4:
        if x.shape[0] != 1:
            raise RuntimeError("eval_crop: one frame at a time (the reference's data has batch 1)")
        x = x[0]
    c, h, w = x.shape
    nh, nw = tile_grid(h, w, patch_size, base)
    out = torch.empty((nh * nw, c, patch_size, patch_size), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().pnnp_eval_crop(x.data_ptr(), out.data_ptr(), c, h, w, patch_size, base, _lib.stream_ptr(x.device)),
                   "eval_crop")
    return out


def eval_merge(croped_data, h, w, base=64):
    """syn_datasets.py:135-159: (nh*nw, c, patch, patch) tiles -> (1, c, h, w): the interior of every tile, later tiles winning."""
    _lib.require_cuda(croped_data, "croped_data")
    t = croped_data.float().contiguous()
    n, c, p, _ = t.shape
    nh, nw = tile_grid(h, w, p, base)
    if n != nh * nw:
        raise RuntimeError(f"eval_merge: {n} tiles for a {nh} x {nw} grid")
    out = torch.empty((1, c, h, w), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().pnnp_eval_merge(t.data_ptr(), out.data_ptr(), c, h, w, p, base, _lib.stream_ptr(t.device)), "eval_merge")
    return out
