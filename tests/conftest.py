import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))   # tests are allowed to import the oracle
GOLDEN = os.path.join(ROOT, "tests", "golden")


# PNNP_GPU_TESTS_ON_CPU_MODELS=1: rehearse `-m gpu` test files without a GPU — every kernel launch goes to the host-compiled device
# source on the CPU models of tests/emul/ (SIMT emulator + tensor-core model), "cuda" means the host (see
# tests/test_gpu_tests_on_cpu_models.py).  Slow for the full-size cases; meant for picking files / -k expressions.
ON_CPU_MODELS = os.environ.get("PNNP_GPU_TESTS_ON_CPU_MODELS") == "1"


@pytest.fixture(autouse=True)
def _gpu_tests_on_cpu_models(request):
    if ON_CPU_MODELS and "gpu" in request.keywords:
        request.getfixturevalue("cuda_is_the_host")
    yield


if ON_CPU_MODELS:
    from test_gpu_tests_on_cpu_models import cuda_is_the_host, libs  # noqa: E402,F401


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    # fp32 references must be true fp32 (cuDNN / cuBLAS default to TF32 for fp32 convs / matmuls on this GPU)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # The driver runs `pytest -x`: the bit-exact / fp32-bound parity tests of the inference path (pack, synthesis, UNet, eval) go
    # first, the tolerance-based training tests (fp32 atomics: results vary with summation order from run to run) go last, so a
    # marginal training tolerance cannot hide the parity results of everything else.  Stable sort: order inside a group is kept.
    def late(it):
        if "test_gpu_wb_jitter.py" in it.nodeid or "test_gpu_preprocess_route.py" in it.nodeid:
            return 2                      # rows added after round 1's GPU budget was spent: rehearsed on the CPU models, first B200 run at round end
        return int("test_gpu_train.py" in it.nodeid or "train_mode" in it.nodeid)
    items.sort(key=late)
    if torch.cuda.is_available() or ON_CPU_MODELS:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def decode_param(d):
    """Inverse of oracle/make_golden.py::_jsonable — restores the exact python / numpy types."""
    out = {}
    for k, v in d.items():
        if "nd" in v:
            out[k] = np.array(v["nd"], dtype=v["dtype"])
        elif "f64" in v:
            out[k] = np.float64(float.fromhex(v["f64"]))
        elif "pyf" in v:
            out[k] = float.fromhex(v["pyf"])
        else:
            out[k] = int(v["pyi"])
    return out


@pytest.fixture(scope="session")
def meta():
    with open(os.path.join(GOLDEN, "meta.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]
    return load
