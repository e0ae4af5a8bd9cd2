"""ctypes binding of the C ABI declared in include/pnnp_b200.h.

The product path has NO CPU fallback: if libpnnp_b200.so is missing, or an entry point
returns non-zero, a RuntimeError is raised (the reference's trainers catch RuntimeError,
trainer_LRID.py:131-135, so the error behaviour of a failed device op is preserved).
"""
import ctypes as C
import os

import torch  # noqa: F401  loads libcudart.so.12 into the process before our library resolves it

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpnnp_b200.so")


class NoiseParamsRow(C.Structure):
    """struct pnnp_noise_params (128 bytes)."""
    _fields_ = [("K", C.c_double), ("sigTL", C.c_double), ("sigGs", C.c_double), ("sigR", C.c_double),
                ("lam", C.c_double), ("q", C.c_double), ("ratio", C.c_double), ("span", C.c_double),
                ("clip_lo", C.c_double), ("bias", C.c_double * 4), ("flags", C.c_uint32),
                ("reserved", C.c_uint32 * 5)]


assert C.sizeof(NoiseParamsRow) == 128


class ConvDesc(C.Structure):
    """struct pnnp_conv_desc."""
    _fields_ = [("mode", C.c_int), ("act", C.c_int), ("out_mode", C.c_int), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("in0", C.c_void_p), ("cin0", C.c_int), ("in1", C.c_void_p), ("cin1", C.c_int),
                ("weight", C.c_void_p), ("w_rows", C.c_int), ("bias", C.c_void_p),
                ("out", C.c_void_p), ("cout", C.c_int), ("cout_stride", C.c_int),
                ("resid", C.c_void_p), ("resid_nchw", C.c_void_p), ("pool_out", C.c_void_p),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_out", C.c_void_p), ("head_cout", C.c_int),
                ("mask", C.c_void_p), ("mask_slope", C.c_float), ("io_f32", C.c_int)]

CODE_P, CODE_G, CODE_R, CODE_Q, CODE_D, CODE_B = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20
CODE_UNIFORM_F64 = 0x100
CHAIN_NUMPY, CHAIN_TORCH = 0, 1
F_K64, F_RATIO64, F_SIG64 = 0x1, 0x2, 0x4

_vp, _i, _u32, _u64, _f, _d = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float, C.c_double

# name -> (restype, argtypes); must list every symbol include/pnnp_b200.h declares
SIGNATURES = {
    "pnnp_last_error": (C.c_char_p, []),
    "pnnp_abi_version": (_i, []),
    "pnnp_launch_count": (_u64, []),
    "pnnp_count_graph_launches": (None, [_u64]),
    "pnnp_pack_norm_u16": (_i, [_vp, _vp, _i, _i, _i, _d, C.POINTER(C.c_double), _i, _i, _vp]),
    "pnnp_pack_norm_f32": (_i, [_vp, _vp, _i, _i, _i, _d, C.POINTER(C.c_double), _i, _i, _vp]),
    "pnnp_pack_norm_dark_u16": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, C.c_double, C.POINTER(C.c_double), _i, _i, C.c_double, _i, C.c_double, _i, _vp]),
    "pnnp_unpack_quant": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _vp]),
    "pnnp_noise_synth": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _u32, _i, _i, _i, _f, _f, _u64, _u64, _u64, _vp]),
    "pnnp_noise_synth_debug": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _u32, _i, _i, _i, _f, _f, _u64, _u64, _u64,
                                    _vp, _vp, _vp, _vp, _vp]),
    "pnnp_noise_synth_replay": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _u32, _i, _i, _i, _f, _f,
                                     _vp, _vp, _vp, _vp, _vp]),
    "pnnp_conv2d_tc": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pnnp_conv2d_tc_ex": (_i, [C.POINTER(ConvDesc), _vp]),
    "pnnp_conv_pipeline_error": (_i, []),
    "pnnp_nchw_to_nhwc16": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "pnnp_conv_first_nchw": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pnnp_conv_first_pipeline_error": (_i, []),
    "pnnp_nchw_to_nhwc16_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "pnnp_maxpool2x2_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "pnnp_l1_loss": (_i, [_vp, _vp, _vp, C.c_size_t, _vp, _vp]),
    "pnnp_head_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pnnp_act_bwd_bias": (_i, [_vp, _vp, _vp, C.c_size_t, _i, _i, _vp]),
    "pnnp_maxpool_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "pnnp_maxpool_bwd_bias": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "pnnp_wgrad_nhwc": (_i, [_i, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp]),
    "pnnp_wgrad_nhwc_pipeline_error": (_i, []),
    "pnnp_adam_step_dev": (_i, [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _f, _f, _f, _f, _vp]),
    "pnnp_strided_copy_batch": (_i, [_vp, _i, _i, _vp]),
    "pnnp_adam_step": (_i, [_vp, _vp, _vp, _vp, C.c_size_t, _f, _f, _f, _f, _i, _f, _vp]),
    "pnnp_crop_aug": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _vp]),
    "pnnp_wb_gains": (_i, [_vp, _i, _i, _i, _i, _f, C.POINTER(C.c_int), C.POINTER(C.c_double), _vp]),
    "pnnp_hbr_map": (_i, [_vp, _vp, C.c_size_t, _vp, _vp, _i, _i, _i, _i, _f, _f, _i, _d, _d, _d, _vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "pnnp_eval_crop": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "pnnp_eval_merge": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "pnnp_eval_epilogue": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp]),
}

class CopyDesc(C.Structure):
    """pnnp_copy_desc (include/pnnp_b200.h)."""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("dst_bf16", C.c_int), ("dim", C.c_int * 4),
                ("sstride", C.c_longlong * 4), ("dstride", C.c_longlong * 4)]


CONV3, CONV1, CONVT, CONV3S2, CONV3X, CONV2S2, CONV3B = 0, 1, 2, 3, 4, 5, 6
ACT_NONE, ACT_LEAKY, ACT_RELU = 0, 1, 2
OUT_NHWC_BF16, OUT_NCHW_F32 = 0, 1

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"pnnp_b200: native library {LIB_PATH} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or pnnp_b200/csrc/build.sh). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"pnnp_b200 {what} failed: {lib().pnnp_last_error().decode(errors='replace')}")


def ptr(t) -> int:
    """Device (or host) pointer of a torch tensor / None."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib().pnnp_launch_count())


def cuda_device(index=None, set_current=False):
    """The CUDA device the product works on: the current one, or `index` (made current when set_current).  No CPU fallback."""
    if not torch.cuda.is_available():
        raise RuntimeError("pnnp_b200: no CUDA device (there is no CPU fallback)")
    if index is None:
        index = torch.cuda.current_device()
    elif set_current:
        torch.cuda.set_device(index)
    return torch.device("cuda", index)


def require_cuda_device(device, what="this operation"):
    if torch.device(device).type != "cuda":
        raise RuntimeError(f"pnnp_b200: {what} needs a CUDA device (no CPU fallback)")


def require_cuda(t, name="tensor"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"pnnp_b200: {name} must be a CUDA tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError(f"pnnp_b200: {name} must be contiguous")
