"""WHOLE warp-collective kernels of the noise synthesis (pnnp_b200/csrc/noise_kernels.cuh) compiled for the host and run, CTA by
CTA, on a lock-step fibre emulator (tests/emul/simt_host.h: every CUDA thread is a fibre; __ballot_sync / __shfl_sync /
__syncwarp / __syncthreads park the fibre until its warp / block has arrived).  The specialised BASELINE-configs[1] kernel —
sampler sorting through the shared-memory queue, shuffled row draws, prefetch, 128-bit accesses — is checked without a GPU:

  * bit-identical output AND draws to the generic kernel (128-bit and scalar paths), for ragged widths, several grid sizes and
    crop offsets: results do not depend on who computes an element;
  * its output is the reference's arithmetic (oracle restatement of process.py:593-631, pinned on the reference goldens) applied to
    the draws it recorded;
  * shard independence: a crop synthesised alone (crop_id0 = k) equals crop k of the batch.

The three MUFU approximations are libm on the host, so the DRAWS differ from a B200's in the last bits; every identity above is
between kernels / the oracle fed the same draws, exactly what the `-m gpu` tests check on the device.  Test infrastructure only."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_np as O
from conftest import ROOT
from pnnp_b200 import _lib
from pnnp_b200.noise import noise_code_bits
from pnnp_b200.noise_params import fill_row

EMUL = os.path.join(ROOT, "tests", "emul")
_f32p, _f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
GENERIC4, GENERIC1, FAST = 0, 1, 2


@pytest.fixture(scope="module")
def S():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(EMUL, "_build", "libsimt_kernels_host.so")
    srcs = [os.path.join(EMUL, f) for f in ("simt_kernels_host.cpp", "simt_host.h", "cuda_host_shim.h")] + \
           [os.path.join(ROOT, "pnnp_b200", "csrc", f) for f in ("noise_kernels.cuh", "noise_core.cuh")] + [os.path.join(ROOT, "include", "pnnp_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-strict-aliasing", "-shared", "-fPIC", "-o", out, srcs[0]], check=True)
    return C.CDLL(out)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _synth(S, clean, params, code, kernel, blocks, chain=_lib.CHAIN_NUMPY, seed=7, offset=3, crop_id0=0, post=(-np.inf, np.inf),
           ori=False, clip=False):
    n, c, h, w = clean.shape
    rows = (_lib.NoiseParamsRow * n)()
    for r, p in zip(rows, params):
        fill_row(r, p, chain == _lib.CHAIN_TORCH)
    clean = np.ascontiguousarray(clean, np.float32)
    out = np.full_like(clean, np.nan)
    d = {"counts": np.full_like(clean, np.nan), "read": np.full_like(clean, np.nan), "row_z": np.full((n, c, h, 1), np.nan, np.float32),
         "q": np.full(clean.shape, np.nan, np.float64)}
    rc = S.emul_noise_synth(_p(clean, _f32p), _p(out, _f32p), rows, n, c, h, w, C.c_uint32(noise_code_bits(code)), chain, int(ori), int(clip),
                            C.c_float(post[0]), C.c_float(post[1]), C.c_uint64(seed), C.c_uint64(offset), C.c_uint64(crop_id0), kernel, 1,
                            _p(d["counts"], _f32p), _p(d["read"], _f32p), _p(d["row_z"], _f32p), _p(d["q"], _f64p), blocks)
    assert rc == 0
    return out, d


def _crops(n, c, h, w, seed):
    rs = np.random.RandomState(seed)
    clean = (rs.rand(n, c, h, w).astype(np.float32)) ** 2
    clean[:, :, :, : w // 8] *= 0.02                                   # a dark stretch: rates below the sampler switch in every row
    np.random.seed(seed)
    return clean, [O.sample_params("SonyA7S2") for _ in range(n)]


@pytest.mark.parametrize("w,blocks,crop_id0", [(512, 1, 0), (644, 3, 5), (1028, 2, 1)])
def test_specialised_kernel_is_bit_identical_to_the_generic_kernel(S, w, blocks, crop_id0):
    clean, params = _crops(3, 4, 3, w, 11 + w)
    fast, df = _synth(S, clean, params, "pgrq", FAST, blocks, crop_id0=crop_id0)
    assert np.isfinite(fast).all()
    for kernel, b in ((GENERIC4, blocks + 1), (GENERIC1, 2)):
        if kernel == GENERIC4 and (crop_id0 * 4 * 3 * w) % 4:
            continue
        gen, dg = _synth(S, clean, params, "pgrq", kernel, b, crop_id0=crop_id0)
        for k in df:
            assert df[k].tobytes() == dg[k].tobytes(), (k, kernel)
        assert fast.tobytes() == gen.tobytes(), kernel
    # both Poisson samplers were exercised, and the counts follow the rates
    lam = clean * np.array([(p["wp"] - p["bl"]) / p["ratio"] / p["K"] for p in params], np.float64).reshape(-1, 1, 1, 1)
    assert (lam < 10).mean() > 0.1 and (lam >= 10).mean() > 0.1
    assert abs((df["counts"] - lam).mean()) < 4 * np.sqrt(lam.mean() / lam.size) + 1e-3


def test_specialised_kernel_output_is_the_reference_arithmetic_on_its_draws(S):
    clean, params = _crops(2, 4, 2, 1152, 5)
    out, d = _synth(S, clean, params, "pgrq", FAST, 2, post=(-np.inf, 1.0))
    for i, p in enumerate(params):
        want = O.noisy_obs_tail_explicit(clean[i], p, "pgrq", {k: v[i] for k, v in d.items()})
        assert np.minimum(want, np.float32(1.0)).tobytes() == out[i].tobytes()            # fused lr.clip(-inf, 1) (syn_datasets.py:339-342)
    assert (out == 1.0).any()
    # row noise: one draw per (crop, channel, row), N(0, 1)-sized; quantisation draws on the 12-bit lattice inside (-0.5, 0.5)
    assert np.abs(d["row_z"]).max() < 6 and len(np.unique(d["row_z"])) == d["row_z"].size
    assert np.abs(d["q"]).max() < 0.5 and np.array_equal(d["q"] * 8192, np.round(d["q"] * 8192))


def test_crops_do_not_depend_on_the_batch_they_are_synthesised_in(S):
    clean, params = _crops(3, 4, 2, 516, 23)
    full, _ = _synth(S, clean, params, "pgrq", FAST, 2, crop_id0=4)
    for k in range(3):
        alone, _ = _synth(S, clean[k:k + 1], params[k:k + 1], "pgrq", FAST, 1, crop_id0=4 + k)
        assert alone[0].tobytes() == full[k].tobytes()
    other, _ = _synth(S, clean, params, "pgrq", FAST, 2, crop_id0=4, offset=4)
    assert other.tobytes() != full.tobytes()                                                # another Philox offset: other draws


@pytest.mark.parametrize("code,chain", [("pgrq", _lib.CHAIN_NUMPY), ("pg", _lib.CHAIN_NUMPY), ("prq", _lib.CHAIN_TORCH), ("gr", _lib.CHAIN_NUMPY)])
def test_generic_kernel_vector_and_scalar_paths_agree_and_follow_the_oracle_tail(S, code, chain):
    clean, params = _crops(2, 4, 2, 260, 31)
    if chain == _lib.CHAIN_TORCH:
        np.random.seed(3)
        params = [O.sample_params_max("SonyA7S2", ratio=None) for _ in range(2)]
    a, da = _synth(S, clean, params, code, GENERIC4, 2, chain=chain)
    b, db = _synth(S, clean, params, code, GENERIC1, 3, chain=chain)
    assert a.tobytes() == b.tobytes()
    if chain == _lib.CHAIN_NUMPY and "p" in code:
        for i, p in enumerate(params):
            draws = {k: v[i] for k, v in da.items()}
            assert O.noisy_obs_tail_explicit(clean[i], p, code, draws).tobytes() == a[i].tobytes()
