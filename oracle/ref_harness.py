"""Live-reference harness (TEST INFRASTRUCTURE — never imported by the product path).

Imports the *unmodified* fenghansen/PNNP sources from /root/reference so that the
restatement in ``oracle/oracle_np.py`` can be validated against them and golden vectors
can be generated (``oracle/make_golden.py``).  /root/reference only exists in the build
container; nothing under ``tests -m gpu`` or ``smoke()`` may call this.  ``bench.py``'s CPU legs
(``--impl reference``, ``cpu_baseline``) load the unmodified copy that ``oracle/build_ref.py`` puts
under the git-ignored ``oracle/_ref/`` (PNNP_REFERENCE_ROOT), never /root/reference itself.

Nine third-party modules the reference imports at module top but never touches on the hot
path are absent from this image; they are replaced by empty stubs.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root():
    """/root/reference in the build container; on the GPU box the unmodified copy under oracle/_ref (oracle/build_ref.py)."""
    if os.path.isdir("/root/reference/data_process"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REFERENCE_ROOT = os.environ.get("PNNP_REFERENCE_ROOT") or _default_root()

_STUBS = {
    "matplotlib": {},
    "matplotlib.pyplot": {},
    "skimage": {},
    "skimage.metrics": {"peak_signal_noise_ratio": None, "structural_similarity": None},
    "exifread": {},
    "rawpy": {},
    "rawpy.enhance": {},
    "h5py": {},
    "natsort": {"natsort": None},
    "torchsummary": {},
}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "data_process"))


def load():
    """Return a namespace with the reference modules: .process, .isp_ops, .archs, .losses, .utils."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch  # noqa: F401  (import before `utils` so its OMP_NUM_THREADS=1 export cannot bind)
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        m.__path__ = []  # behave like a package for "a.b" imports
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import utils as ref_utils          # noqa: E402  (seeds everything to 1997, utils/utils.py:45-48)
    import data_process               # noqa: E402,F401
    import archs as ref_archs          # noqa: E402
    import losses as ref_losses        # noqa: E402
    ns = types.SimpleNamespace()
    ns.utils = ref_utils
    # data_process/__init__.py star-imports a *function* named `process`, shadowing the submodule
    ns.process = sys.modules["data_process.process"]
    ns.isp_ops = sys.modules["utils.isp_ops"]
    ns.syn_datasets = sys.modules["data_process.syn_datasets"]
    ns.data_process = sys.modules["data_process"]
    ns.archs = ref_archs
    ns.losses = ref_losses
    return ns


def versions() -> dict:
    import numpy, scipy, torch
    return {"numpy": numpy.__version__, "scipy": scipy.__version__, "torch": torch.__version__}
