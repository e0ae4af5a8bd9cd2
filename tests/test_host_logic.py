"""Product host logic that needs no GPU: parameter sampling, table packing, the C-ABI surface."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

import pnnp_b200 as P
from pnnp_b200 import _lib
from pnnp_b200.noise_params import fill_row
from conftest import ROOT, decode_param


def test_param_sampling_matches_reference_goldens(meta):
    for case in meta["params"]:
        np.random.seed(case["seed"])
        got = getattr(P, case["fn"])(case["camera"], **case["kwargs"])
        want = decode_param(case["out"])
        assert got.keys() == want.keys()
        for k in want:
            assert type(got[k]) is type(want[k]), (case["fn"], case["camera"], case["kwargs"], k)
            assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), (case, k)


def test_param_sampling_reference_errors():
    for cam in ("IMX686", "NikonD850"):
        with pytest.raises(KeyError, match="uReadk"):
            P.sample_params(cam)
    with pytest.raises(KeyError):
        P.sample_params_max("SonyA7S2", iso=123)


def test_noise_code_bits():
    assert P.noise_code_bits("pgrq") == 0x0F and P.noise_code_bits("PR") == 0x05
    assert P.noise_code_bits("pgrqdb") == 0x3F and P.noise_code_bits("") == 0


def test_param_row_flags():
    np.random.seed(0)
    row = _lib.NoiseParamsRow()
    fill_row(row, P.sample_params("SonyA7S2"))
    assert row.flags == (_lib.F_K64 | _lib.F_SIG64) and row.span == 15871 and row.clip_lo == -512 / 16383
    fill_row(row, P.sample_params_max("SonyA7S2", iso=1600))
    assert row.flags == 0
    fill_row(row, P.sample_params_max("IMX686", iso=6400))
    assert row.flags == _lib.F_RATIO64 and list(row.bias) == [-0.08113494, -0.04906388, -0.9408157, -1.2048522]
    fill_row(row, P.sample_params_max("IMX686"), torch_chain=True)
    assert row.flags == 0 and row.clip_lo == float(np.float32(-64.0) / np.float32(1023.0))


def test_error_mirroring_without_gpu():
    p = {"K": 1.0, "sigTL": 1.0, "sigR": 1.0, "sigGs": 1.0, "bias": 0, "lam": 0.1, "q": 1e-3, "ratio": 2.0,
         "wp": 1023, "bl": 64}
    with pytest.raises(AttributeError):
        P.generate_noisy_obs(np.zeros((4, 2, 2), np.float32), param=p, noise_code="pd")
    import torch
    t = torch.zeros(4, 2, 2)
    with pytest.raises(TypeError):
        P.generate_noisy_torch(t, param=p, noise_code="r")
    with pytest.raises(NotImplementedError):
        P.generate_noisy_torch(t, param=p, noise_code="pg")
    with pytest.raises(TypeError):
        P.generate_noisy_torch(t, param=p, noise_code="pd")


def test_abi_exports_every_declared_symbol():
    """The shared library loads and exports exactly what include/pnnp_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "pnnp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pnnp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 9
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().pnnp_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA tensor"):
        P.generate_noisy_obs(np.zeros((4, 4, 4), np.float32),
                             param={"K": 1.0, "sigTL": 1.0, "sigR": 1.0, "sigGs": 1.0, "bias": 0, "lam": 0.1,
                                    "q": 1e-3, "ratio": 2.0, "wp": 1023, "bl": 64}, noise_code="p")
    with pytest.raises(RuntimeError):
        P.raw2bayer(np.zeros((4, 4), np.uint16))


def test_host_emulation_is_unreachable_from_the_product(tmp_path, monkeypatch):
    """The CPU models of tests/emul/ are test infrastructure: the product build never defines PNNP_HOST_EMUL (the only way the csrc
    headers see tests/emul/), no Python module of the package refers to tests/ or to an emulated library, the built library exports
    no emulation entry point, and without a CUDA device the trainer, the dataset and the training step raise."""
    import subprocess
    import torch
    pkg = os.path.join(ROOT, "pnnp_b200")
    assert "PNNP_HOST_EMUL" not in open(os.path.join(pkg, "csrc", "build.sh")).read()
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "emul" not in src.lower() and "tests/emul" not in src and "cuda_is_the_host" not in src, fn
    so = os.path.join(pkg, "libpnnp_b200.so")
    if os.path.exists(so):
        syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
        assert "emul_" not in syms and "simt" not in syms
    if torch.cuda.is_available():
        return
    import yaml
    from pnnp_b200 import trainer as T
    from pnnp_b200.datasets import Raw_Dataset
    from pnnp_b200.train import UNetTrainStep
    cfg = yaml.load(open(os.path.join(ROOT, "runfiles/SonyA7S2/PNNP.yml")), Loader=yaml.FullLoader)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Raw_Dataset(dict(cfg["dst_train"], H=64, W=64, patch_size=32))[0]
    with pytest.raises(RuntimeError, match="CUDA device"):
        UNetTrainStep(P.UNetSeeInDark(cfg["arch"]))
    monkeypatch.chdir(tmp_path)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        T.SID_Trainer(["-f", os.path.join(ROOT, "runfiles/SonyA7S2/PNNP.yml"), "--mode", "train"])


def test_product_never_imports_oracle():
    """No import / include / dlopen of anything under oracle/ from the product package."""
    pkg = os.path.join(ROOT, "pnnp_b200")
    pat_py = re.compile(r"^\s*(import|from)\s+(oracle|oracle_np|ref_harness)\b|sys\.path.*oracle|CDLL\(.*oracle", re.M)
    pat_c = re.compile(r"#\s*include\s*[\"<][^\">]*oracle")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            src_path = os.path.join(dp, fn)
            if fn.endswith(".py"):
                assert not pat_py.search(open(src_path).read()), fn
            elif fn.endswith((".cu", ".cuh", ".h", ".sh")):
                assert not pat_c.search(open(src_path).read()), fn


def test_lr_schedules_match_the_reference_functions():
    """base_trainer.py:33-43,141-160: WarmupCosine / MultiStep as the YAML `hyper` block selects them."""
    import math
    from pnnp_b200.utils import get_cos_lr, get_multistep_lr, lr_lambda_from_hyper
    hyper = {"lr_scheduler": "WarmupCosine", "learning_rate": 1e-4, "last_epoch": 0, "stop_epoch": 1600, "step_size": 10, "T": 2}
    f = lr_lambda_from_hyper(hyper)                                      # period 800, peak 10
    assert f(10) == pytest.approx(1e-4) and f(800 - 1e-9) == pytest.approx(0.2e-4, rel=1e-3)
    assert f(800) == 0.0 and f(805) == pytest.approx(0.5 * 1e-4 * 0.5) and f(810) == pytest.approx(0.5e-4)   # warm-up of period 2
    assert f(1) == pytest.approx(1e-4 * (0.8 * (math.cos((1 - 10) / 790 * math.pi) * 0.5 + 0.5) + 0.2))
    g = lr_lambda_from_hyper(dict(hyper, lr_scheduler="MultiStep", step_size=100))
    assert [g(e) for e in (1, 100, 101, 180, 181, 799, 900, 901)] == pytest.approx([1e-4, 1e-4, 5e-5, 5e-5, 1e-5, 1e-5, 1e-4, 5e-5])
    with pytest.raises(KeyError):
        lr_lambda_from_hyper(dict(hyper, lr_scheduler="linear"))
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):                                          # live check against the unmodified reference
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_harness as rh
        rh.load()
        B = sys.modules.get("base_trainer") or __import__("base_trainer")
        for step in (0, 1, 9, 10, 11, 400, 799, 800, 801, 805, 1599):
            assert get_cos_lr(step, period=800, peak=10, lr=1e-4) == B.get_cos_lr(step, period=800, peak=10, lr=1e-4)
            assert get_multistep_lr(step, period=800, lr=1e-4, milestone=[10, 18]) == B.get_multistep_lr(step, period=800, lr=1e-4, milestone=[10, 18])


def test_built_library_uses_tcgen05_tma_and_packed_fp32():
    """The machine code in pnnp_b200/libpnnp_b200.so (sm_100a): the conv / wgrad kernels issue tcgen05.mma (SASS UTCHMMA) with
    operands staged by TMA (UTMALDG) and accumulators read back from TMEM (LDTM); no legacy mma.sync (HMMA) anywhere; the noise
    kernel streams with 128-bit accesses; the opt-in epilogues use packed fp32 pairs (FADD2 / FFMA2)."""
    import shutil
    import subprocess
    from pnnp_b200 import _lib
    if shutil.which("cuobjdump") is None or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "arch = sm_100a" in sass
    per_fn, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            per_fn[cur] = []
        elif cur is not None:
            per_fn[cur].append(line)
    conv = {k: "\n".join(v) for k, v in per_fn.items() if "conv_gemm_tc_kernel" in k}
    wgrad = {k: "\n".join(v) for k, v in per_fn.items() if "wgrad_nhwc_kernel" in k}
    assert len(conv) >= 27 and len(wgrad) >= 1
    for body in list(conv.values()) + list(wgrad.values()):
        assert "UTCHMMA" in body and "UTMALDG" in body and "LDTM" in body and "UTCBAR" in body
    assert "HMMA." not in sass and "IMMA." not in sass
    noise = "\n".join(per_fn[next(k for k in per_fn if "noise_synth_fast_kernelILb0" in k)])
    import re
    assert re.search(r"LDG\.E\S*\.128", noise) and re.search(r"STG\.E\S*\.128", noise)      # 128-bit coalesced HBM loads and stores
    packed = [k for k in conv if k.endswith("ELi4EEEv14CUtensorMap_stS1_S1_NS_10ConvParamsE")]
    assert packed and all("FADD2" in conv[k] and "FFMA2" in conv[k] for k in packed)


def test_bayer2rggb_rggb2bayer_true_rggb_order_and_round_trip():
    """P3 (utils/isp_ops.py:57-63): HWC pack in TRUE RGGB order — R (0,0), G1 (0,1), G2 (1,0), B (1,1) — which differs from
    raw2bayer's plane order R, G1, B, G2 (isp_ops.py:87-90); rggb2bayer is its inverse.  Checked against the index definition,
    the oracle, and (build container) the live reference."""
    import oracle_np as O
    import ref_harness as rh
    rs = np.random.RandomState(11)
    for H, W, dt in ((4, 6, np.uint16), (64, 96, np.float32), (2, 2, np.int32)):
        bayer = (rs.rand(H, W) * 16383).astype(dt)
        rggb = P.bayer2rggb(bayer)
        assert rggb.shape == (H // 2, W // 2, 4) and rggb.dtype == dt
        for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            assert np.array_equal(rggb[..., k], bayer[dy::2, dx::2])
        assert np.array_equal(P.rggb2bayer(rggb), bayer)
        assert np.array_equal(rggb, O.bayer2rggb(bayer)) and np.array_equal(O.rggb2bayer(rggb), bayer)
        if rh.available():
            ISP = rh.load().isp_ops
            assert np.array_equal(rggb, ISP.bayer2rggb(bayer)) and np.array_equal(P.rggb2bayer(rggb), ISP.rggb2bayer(rggb))
    # not the same channel order as the packed planes raw2bayer produces: planes 2 and 3 are swapped
    bayer = np.arange(16, dtype=np.float32).reshape(4, 4)
    planes = O.raw2bayer(bayer, wp=1, bl=0, norm=False)
    assert np.array_equal(planes[[0, 1, 3, 2]].transpose(1, 2, 0), P.bayer2rggb(bayer))


def test_lr_of_the_kth_trained_epoch_matches_the_reference_scheduler_with_nonzero_last_epoch():
    """trainer_SID.py:57,75,127 + base_trainer.py:131-138: LambdaScheduler (a LambdaLR whose get_lr returns lmbda(last_epoch)
    itself) starts at -1 whatever hyper['last_epoch'] is.  With the reference PNNP.yml's resume settings (last_epoch 1200,
    stop_epoch 1600, T 2) the first trained epoch must run at get_cos_lr(1), not get_cos_lr(1201) (cycle 6, lr / 64)."""
    import torch
    from torch.optim.lr_scheduler import LambdaLR
    from pnnp_b200.utils import lr_for_epoch, lr_lambda_from_hyper

    class LambdaScheduler(LambdaLR):                       # the reference's class, restated for the test
        def get_lr(self):
            return [lmbda(self.last_epoch) for lmbda, _ in zip(self.lr_lambdas, self.base_lrs)]

    for last_epoch, stop_epoch, sched in ((1200, 1600, "WarmupCosine"), (0, 1600, "WarmupCosine"), (300, 700, "MultiStep")):
        hyper = {"lr_scheduler": sched, "learning_rate": 1e-4, "last_epoch": last_epoch, "stop_epoch": stop_epoch,
                 "step_size": 10, "T": 2}
        lam = lr_lambda_from_hyper(hyper)
        opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=hyper["learning_rate"])
        sch = LambdaScheduler(opt, lam)
        sch.step()                                         # top of train()
        for epoch in range(last_epoch + 1, last_epoch + 260):
            assert lr_for_epoch(lam, epoch, last_epoch) == sch.get_last_lr()[0], (sched, epoch)
            opt.step()
            sch.step()                                     # after each trained epoch
    hyper = {"lr_scheduler": "WarmupCosine", "learning_rate": 1e-4, "last_epoch": 1200, "stop_epoch": 1600, "step_size": 10, "T": 2}
    lam = lr_lambda_from_hyper(hyper)
    assert lr_for_epoch(lam, 1201, 1200) == lam(1) and lam(1) > 32 * lam(1201)


def test_reference_copy_recipe_is_verbatim_and_the_bench_baseline_uses_it(tmp_path):
    """oracle/build_ref.py copies the reference packages byte for byte into a (git-ignored) directory with a SHA-256 manifest;
    bench.py's CPU legs run that copy (`cpu_baseline.kind == "reference"`) and fall back to the oracle port when it is absent or
    has been touched."""
    import build_ref
    src = "/root/reference"
    if not os.path.isdir(os.path.join(src, "data_process")):
        pytest.skip("reference tree not present")
    dest = build_ref.build(src, str(tmp_path / "_ref"), quiet=True)
    assert build_ref.verify(dest)
    n = 0
    for root, _d, files in os.walk(dest):
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, f), dest)
                assert open(os.path.join(root, f), "rb").read() == open(os.path.join(src, rel), "rb").read(), rel
                n += 1
    assert n >= 20
    with open(os.path.join(dest, "archs", "Unet.py"), "a") as f:
        f.write("\n# edited\n")
    assert not build_ref.verify(dest)                                   # an edited copy is refused
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in gi                                         # never part of the history
    ig = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(ig) or "oracle/_ref" not in open(ig).read()   # but it travels to the GPU box
