"""Physics-based noise synthesis behind the reference's call signatures
(data_process/process.py:591-673), executed by ONE fused sm_100a kernel (csrc/noise_synth.cu).

  generate_noisy_obs   — NumPy-chain arithmetic (float64 where NEP-50 promotes), the function the
                         DataLoader workers call (syn_datasets.py:326-337);
  generate_noisy_torch — float32-chain arithmetic, the function trainer_SID.py:449-462 calls;
  synthesize_batch     — the batched entry both shims use: all crops of a step in one launch.

The draws come from Philox4x32-10 keyed on (seed, crop, element); `replay_*` feeds the
reference's own draws through the same arithmetic core (bit-exact parity tests).
"""
import math

import numpy as np
import torch

from . import _lib
from .noise_params import ParamTable
from .rng import default_generator

_LETTERS = (("p", _lib.CODE_P), ("g", _lib.CODE_G), ("r", _lib.CODE_R), ("q", _lib.CODE_Q),
            ("d", _lib.CODE_D), ("b", _lib.CODE_B))


def noise_code_bits(noise_code: str) -> int:
    """process.py:598-603 — substring tests on the lower-cased code."""
    c = noise_code.lower()
    return sum(bit for ch, bit in _LETTERS if ch in c)


def _as_batch(y):
    """(c,h,w) or (n,c,h,w) → contiguous float32 CUDA (n,c,h,w) + flag."""
    single = y.dim() == 3
    t = y.unsqueeze(0) if single else y
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t, single


def synthesize_batch(clean, params, noise_code="p", chain=_lib.CHAIN_NUMPY, ori=False, clip=False,
                     post_clip=None, generator=None, crop_id0=0, out=None, table=None, debug=False,
                     table_row0=0, seed_offset=None):
    """clean: CUDA float32 (n,c,h,w); params: list of n reference-style dicts (or a ParamTable via
    `table`).  post_clip=(lo, hi) fuses the caller's follow-up clamp (syn_datasets.py:339-342).
    Returns noisy (and a dict of the draws when debug=True)."""
    _lib.require_cuda(clean, "clean")
    n, c, h, w = clean.shape
    if table is None:
        table = ParamTable(params, clean.device, torch_chain=(chain == _lib.CHAIN_TORCH))
    if table.n < table_row0 + n:
        raise RuntimeError(f"pnnp_b200: {table.n} parameter rows for crops [{table_row0}, {table_row0 + n})")
    if out is None:
        out = torch.empty_like(clean)
    if seed_offset is None:
        gen = default_generator if generator is None else generator
        seed, offset = gen.next()
    else:
        seed, offset = seed_offset
    tab_ptr = table.data_ptr() + 128 * table_row0
    lo, hi = (-math.inf, math.inf) if post_clip is None else post_clip
    bits = noise_code_bits(noise_code) if isinstance(noise_code, str) else int(noise_code)
    if chain == _lib.CHAIN_NUMPY and getattr(table, "uniform_f64", False):
        bits |= _lib.CODE_UNIFORM_F64
    L = _lib.lib()
    with torch.cuda.device(clean.device):
        st = _lib.stream_ptr(clean.device)
        if not debug:
            _lib.check(L.pnnp_noise_synth(clean.data_ptr(), out.data_ptr(), tab_ptr, n, c, h, w, bits, chain,
                                          int(bool(ori)), int(bool(clip)), lo, hi, seed, offset, crop_id0, st),
                       "noise_synth")
            return out
        d = {"shot": torch.empty_like(clean), "read": torch.empty_like(clean),
             "row_z": torch.zeros((n, c, h, 1), dtype=torch.float32, device=clean.device),
             "q": torch.empty(clean.shape, dtype=torch.float64, device=clean.device)}
        _lib.check(L.pnnp_noise_synth_debug(clean.data_ptr(), out.data_ptr(), tab_ptr, n, c, h, w, bits, chain,
                                            int(bool(ori)), int(bool(clip)), lo, hi, seed, offset, crop_id0,
                                            d["shot"].data_ptr(), d["read"].data_ptr(), d["row_z"].data_ptr(),
                                            d["q"].data_ptr(), st), "noise_synth_debug")
        return out, d


def replay_batch(clean, params, noise_code, draws, chain=_lib.CHAIN_NUMPY, ori=False, clip=False, post_clip=None):
    """Same arithmetic core with caller-supplied draws: shot (counts | N(0,1)), read (DN),
    row_z (n,c,h,1), q (float64: numpy chain U(-.5,.5) DN, torch chain U[0,1))."""
    _lib.require_cuda(clean, "clean")
    n, c, h, w = clean.shape
    table = ParamTable(params, clean.device, torch_chain=(chain == _lib.CHAIN_TORCH))
    out = torch.empty_like(clean)
    dev = clean.device

    def prep(key, dtype):
        v = draws.get(key)
        if v is None:
            return None
        t = torch.as_tensor(np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v)
        return t.to(device=dev, dtype=dtype).contiguous()

    shot, read = prep("shot", torch.float32), prep("read", torch.float32)
    rowz, q = prep("row_z", torch.float32), prep("q", torch.float64)
    lo, hi = (-math.inf, math.inf) if post_clip is None else post_clip
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pnnp_noise_synth_replay(
            clean.data_ptr(), out.data_ptr(), table.data_ptr(), n, c, h, w, noise_code_bits(noise_code), chain,
            int(bool(ori)), int(bool(clip)), lo, hi, _lib.ptr(shot), _lib.ptr(read), _lib.ptr(rowz), _lib.ptr(q),
            _lib.stream_ptr(dev)), "noise_synth_replay")
    return out


def _check_python_errors(noise_code, param, torch_chain):
    """Reproduce the exceptions the reference raises for unsupported combinations."""
    c = noise_code.lower()
    if torch_chain:
        if "p" not in c:
            raise TypeError("generate_noisy_torch needs 'p' in noise_code (reference: process.py:651)")
        if "g" in c and "b" not in c:
            raise NotImplementedError  # process.py:654
        if "d" in c:
            raise TypeError("generate_noisy_torch does not support 'd' (reference: process.py:663)")
    elif "d" in c and "b" not in c and not hasattr(param["bias"], "reshape"):
        raise AttributeError(f"'{type(param['bias']).__name__}' object has no attribute 'reshape'")  # process.py:617


def generate_noisy_obs(y, camera_type=None, wp=16383, noise_code="p", param=None, MultiFrameMean=1, ori=False,
                       clip=False):
    """process.py:591-631.  NumPy array in → float32 NumPy array out (H2D + kernel + D2H);
    CUDA tensor in → CUDA tensor out."""
    if MultiFrameMean != 1:
        raise NotImplementedError("MultiFrameMean != 1 is not used by any reference call site")
    _check_python_errors(noise_code, param, torch_chain=False)
    host = not isinstance(y, torch.Tensor)
    if host:
        if not torch.cuda.is_available():
            raise RuntimeError("pnnp_b200: no CUDA device (there is no CPU fallback)")
        t = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).cuda()
    else:
        t = y
    t, single = _as_batch(t)
    plist = [param] * t.shape[0]
    out = synthesize_batch(t, plist, noise_code, _lib.CHAIN_NUMPY, ori, clip)
    out = out[0] if single else out
    return out.cpu().numpy() if host else out


def generate_noisy_torch(y, camera_type=None, noise_code="p", param=None, MultiFrameMean=1, ori=False, clip=False):
    """process.py:634-673.  y: CUDA tensor (c,h,w) or (n,c,h,w); param values may be 0-d tensors
    (what trainer_SID.py:455-458 passes) or python numbers."""
    if MultiFrameMean != 1:
        raise NotImplementedError("MultiFrameMean != 1 is not used by any reference call site")
    _check_python_errors(noise_code, param, torch_chain=True)
    t, single = _as_batch(y)
    out = synthesize_batch(t, [param] * t.shape[0], noise_code, _lib.CHAIN_TORCH, ori, clip)
    return out[0] if single else out
