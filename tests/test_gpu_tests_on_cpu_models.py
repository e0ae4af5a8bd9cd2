"""`-m gpu` test functions that had not run on a B200 when round 1's GPU budget ended (white-balance jitter, the trainer's
`gpu_preprocess` route, the LRID eval entry point's reflect-pad branch) executed HERE, unchanged, on the CPU models: the emulated
library of tests/test_device_tc_on_cpu.py stands where libpnnp_b200.so stands, "cuda" means the host.  They still run on the
device in the `-m gpu` suite; this is the rehearsal.  Test infrastructure only."""
import contextlib
import functools

import pytest
import torch

from pnnp_b200 import _lib
from test_device_tc_on_cpu import _EmulatedLibrary, _VARIANT_ENV, libs  # noqa: F401  (libs: fixture)


def _on_host(fn):
    """torch factory called with device='cuda' -> the host."""
    @functools.wraps(fn)
    def wrapped(*a, **k):
        if "device" in k and str(k["device"]).startswith("cuda"):
            k["device"] = "cpu"
        return fn(*a, **k)
    return wrapped


_Generator = torch.Generator


class _HostGenerator(_Generator):
    """torch.Generator(device='cuda') -> a host generator (a subclass: annotations such as `torch.Generator | None` keep working)."""

    def __new__(cls, device="cpu"):
        return _Generator.__new__(cls, device="cpu" if str(device).startswith("cuda") else device)


def _to_host(orig):
    @functools.wraps(orig)
    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) or (isinstance(x, torch.device) and x.type == "cuda") else x for x in a)
        if "device" in k and str(k["device"]).startswith("cuda"):
            k["device"] = "cpu"
        return orig(self, *a, **k)
    return to


@pytest.fixture
def cuda_is_the_host(monkeypatch, libs):  # noqa: F811
    lib = _EmulatedLibrary(*libs)
    monkeypatch.setattr(_lib, "lib", lambda: lib)
    monkeypatch.setattr(_lib, "stream_ptr", lambda device=None: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name="tensor": None)
    monkeypatch.setattr(_lib, "require_cuda_device", lambda device, what="": None)
    monkeypatch.setattr(_lib, "cuda_device", lambda index=None, set_current=False: torch.device("cpu"))
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda i: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)        # torch.optim asks before every step
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self.clone())     # a copy, as a host-to-device transfer is
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    for name in ("rand", "randn", "zeros", "ones", "empty", "full", "arange", "tensor", "as_tensor", "zeros_like", "empty_like", "randint"):
        monkeypatch.setattr(torch, name, _on_host(getattr(torch, name)))
    monkeypatch.setattr(torch, "Generator", _HostGenerator)
    monkeypatch.setattr(torch.Tensor, "to", _to_host(torch.Tensor.to))
    monkeypatch.setattr(torch.nn.Module, "to", _to_host(torch.nn.Module.to))
    monkeypatch.setenv("PNNP_TRAIN_GRAPH", "0")
    for k in _VARIANT_ENV:
        monkeypatch.delenv(k, raising=False)
    yield lib
    assert lib.pnnp_conv_pipeline_error() == 0 and lib.pnnp_wgrad_nhwc_pipeline_error() == 0


@pytest.mark.parametrize("tag", ["wb32", "wb64", "wbpy"])
def test_wb_jitter_golden(cuda_is_the_host, golden, tag):
    import test_gpu_wb_jitter as W
    W.test_wb_gains_kernel_is_bit_exact_vs_reference_golden(golden, tag)


def test_wb_jitter_oracle_and_dataset_order(cuda_is_the_host):
    import test_gpu_wb_jitter as W
    W.test_wb_gains_kernel_vs_oracle_on_crops()
    W.test_raw_dataset_item_with_wb_jitter_follows_the_reference_order()


def test_gpu_preprocess_route(cuda_is_the_host):
    import test_gpu_preprocess_route as R
    R.test_raw_dataset_leaves_the_noise_to_the_trainer()
    R.test_preprocess_train_is_the_reference_loop_in_one_launch()


def test_lrid_eval_entry_point_reflect_pad_branch(cuda_is_the_host, tmp_path, monkeypatch):
    import test_gpu_trainer as G
    G.test_lrid_eval_entry_point_takes_reflect_pad_branch(tmp_path, monkeypatch)


@pytest.mark.parametrize("shape,scale,correct", [((1, 4, 64, 96), 1.0, False), ((2, 3, 45, 70), 1.5, False), ((1, 4, 128, 192), 1.0, True)])
def test_experimental_separable_ssim(cuda_is_the_host, monkeypatch, shape, scale, correct):
    import test_gpu_variants as E
    E.test_separable_ssim_equals_default_kernel(monkeypatch, shape, scale, correct)


@pytest.mark.parametrize("act_kind", [0, 1, 2])
def test_experimental_act_backward_v2(cuda_is_the_host, monkeypatch, act_kind):
    import test_gpu_variants as E
    E.test_act_backward_v2_equals_the_default_kernel(monkeypatch, act_kind)


def test_experimental_wgrad_v2(cuda_is_the_host, monkeypatch):
    import test_gpu_variants as E
    E.test_wgrad_v2_matches_autograd_like_the_default_kernel(monkeypatch)


def test_resunet_backward_on_the_cpu_models(cuda_is_the_host):
    """train_resunet.ResUnetTrainStep (round 2): residual blocks, 1x1 shortcuts, stride-2 convs through zero-inserted gradients —
    the product's host code with every launch on the CPU models, against fp32 autograd through the oracle's functional ResUnet."""
    import test_gpu_train_resunet as R
    R.test_resunet_backward_matches_fp32_autograd()
