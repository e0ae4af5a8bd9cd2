"""The DEVICE source of the noise synthesis core (pnnp_b200/csrc/noise_core.cuh: Philox, the samplers, the deterministic tails)
compiled for the host through tests/emul/cuda_host_shim.h and run on the CPU:

  * the tails on the reference's own draws  -> bit-exact against the goldens of the unmodified reference (the same check the
    `-m gpu` replay tests make on the device, here on the very same source lines without a GPU);
  * the specialised kernel's Markstein divisions -> equal to IEEE division, and its arithmetic equal to the generic tail;
  * Philox blocks -> equal to the CPU Philox of the oracle;
  * the samplers (normal, Tukey-lambda, Poisson below / above the switch) -> KS / chi-square against the exact distributions
    (the three MUFU approximations are libm here: sampler parity is statistical on the device as well).

Test infrastructure only: nothing in the product can reach this code (the library has no CPU path)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
from scipy import stats

import oracle_np as O
from conftest import ROOT, decode_param
from pnnp_b200 import _lib
from pnnp_b200.noise import noise_code_bits
from pnnp_b200.noise_params import fill_row

EMUL = os.path.join(ROOT, "tests", "emul")
_f32p, _f64p, _u32p, _u64p = (C.POINTER(t) for t in (C.c_float, C.c_double, C.c_uint32, C.c_uint64))


@pytest.fixture(scope="module")
def core():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(EMUL, "_build", "libnoise_core_host.so")
    srcs = [os.path.join(EMUL, "noise_core_host.cpp"), os.path.join(EMUL, "cuda_host_shim.h"),
            os.path.join(ROOT, "pnnp_b200", "csrc", "noise_core.cuh"), os.path.join(ROOT, "pnnp_b200", "csrc", "pack_core.cuh"),
            os.path.join(ROOT, "include", "pnnp_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-strict-aliasing", "-shared", "-fPIC", "-o", out, srcs[0]], check=True)
    return C.CDLL(out)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _row(p, torch_chain=False):
    r = _lib.NoiseParamsRow()
    fill_row(r, p, torch_chain)
    return r


def _replay(core, y, p, code, chain, ori, clip, shot=None, read=None, row_z=None, q=None):
    c, h, w = y.shape
    y = np.ascontiguousarray(y, np.float32)
    out = np.empty_like(y)
    shot = None if shot is None else np.ascontiguousarray(shot, np.float32)
    read = None if read is None else np.ascontiguousarray(read, np.float32)
    row_z = None if row_z is None else np.ascontiguousarray(row_z, np.float32).reshape(c * h)
    q = None if q is None else np.ascontiguousarray(q, np.float64)
    r = _row(p, chain == _lib.CHAIN_TORCH)
    core.emul_replay(_p(y, _f32p), _p(out, _f32p), C.byref(r), c, h, w, C.c_uint32(noise_code_bits(code)), chain, int(bool(ori)),
                     int(bool(clip)), C.c_float(-np.inf), C.c_float(np.inf), _p(shot, _f32p), _p(read, _f32p), _p(row_z, _f32p),
                     _p(q, _f64p))
    return out


def test_tail_numpy_source_is_bit_exact_vs_reference_goldens(core, golden, meta):
    g = golden("noisy_obs")
    y = g["y"]
    for c in meta["noisy_obs_cases"]:
        t = c["tag"]
        shot = g[t + "_counts"] if t + "_counts" in g.files else (g[t + "_shot_z"] if t + "_shot_z" in g.files else None)
        d = {k: (g[f"{t}_{k}"] if f"{t}_{k}" in g.files else None) for k in ("read", "row_z", "q")}
        out = _replay(core, y, decode_param(c["param"]), c["code"], _lib.CHAIN_NUMPY, c["ori"], c["clip"], shot, **d)
        assert out.tobytes() == g[t + "_z"].tobytes(), c


def test_tail_torch_source_is_bit_exact_vs_reference_goldens(core, golden, meta):
    g = golden("noisy_torch")
    y = g["y"]
    for c in meta["noisy_torch_cases"]:
        t = c["tag"]
        out = _replay(core, y, decode_param(c["param"]), c["code"], _lib.CHAIN_TORCH, c["ori"], bool(c["clip"]),
                      g[t + "_counts"], g[t + "_read"], g[t + "_row_z"] if t + "_row_z" in g.files else None,
                      g[t + "_q_u"].astype(np.float64) if t + "_q_u" in g.files else None)
        assert out.tobytes() == g[t + "_z"].tobytes(), c


def test_markstein_division_equals_ieee_division(core):
    rs = np.random.RandomState(0)
    n = 2_000_000
    a32 = (rs.standard_normal(n) * 10.0 ** rs.uniform(-6, 6, n)).astype(np.float32)
    b32 = (rs.uniform(0.5, 2.0, n) * 10.0 ** rs.uniform(-3, 5, n)).astype(np.float32)
    out32 = np.empty_like(a32)
    core.emul_div_by_const_f32(_p(a32, _f32p), _p(b32, _f32p), n, _p(out32, _f32p))
    assert np.array_equal(out32, a32 / b32)
    a64 = rs.standard_normal(n) * 10.0 ** rs.uniform(-8, 8, n)
    b64 = rs.uniform(0.5, 2.0, n) * 10.0 ** rs.uniform(-3, 6, n)
    b64[:4] = [15871.0, 959.0, 16383.0 - 512.0, 1023.0 - 64.0]                  # the spans the kernel divides by
    out64 = np.empty_like(a64)
    core.emul_div_by_const_f64(_p(a64, _f64p), _p(b64, _f64p), n, _p(out64, _f64p))
    assert np.array_equal(out64, a64 / b64)


def test_specialised_kernel_arithmetic_equals_the_generic_tail_and_the_oracle(core):
    """'pgrq' with sample_params output (np.float64 K / sigR, python-float ratio): the fast kernel's reciprocal-multiply chain ==
    tail_numpy == the oracle's explicit restatement, and its rate == the generic rate, on a crop row of real draws."""
    rs = np.random.RandomState(2)
    for seed in range(6):
        np.random.seed(seed)
        p = O.sample_params("SonyA7S2")
        y = (rs.rand(1, 1, 4096).astype(np.float32)) ** 2
        np.random.seed(100 + seed)
        want, d = O.generate_noisy_obs(y, param=p, noise_code="pgrq", return_draws=True)
        assert O.noisy_obs_tail_explicit(y, p, "pgrq", d).tobytes() == want.tobytes()
        generic = _replay(core, y, p, "pgrq", _lib.CHAIN_NUMPY, False, False, d["counts"], d["read"], d["row_z"], d["q"])
        assert generic.tobytes() == want.tobytes()
        r = _row(p)
        cnt, read, q = (np.ascontiguousarray(d[k].reshape(-1), t) for k, t in (("counts", np.float32), ("read", np.float32), ("q", np.float64)))
        fast, rate = np.empty(4096, np.float32), np.empty(4096, np.float32)
        core.emul_fast_tail(_p(np.ascontiguousarray(y.reshape(-1)), _f32p), _p(fast, _f32p), _p(rate, _f32p), C.byref(r), 4096,
                            _p(cnt, _f32p), _p(read, _f32p), C.c_float(float(d["row_z"].reshape(-1)[0])), _p(q, _f64p),
                            C.c_float(-np.inf), C.c_float(np.inf))
        assert fast.tobytes() == want.tobytes()
        # the specialised kernel folds np.clip(z, lo, 1) * ratio and the caller's post-clip into ONE float32 clamp (noise_core.cuh:
        # fast_constants): equal to clipping the reference's output, also where the two clips cut into each other
        for lo, hi in ((-np.inf, 1.0), (0.0, 0.5), (-0.01, 250.0), (float(want.max()) * 2, np.inf), (-np.inf, float(want.min()) - 1)):
            core.emul_fast_tail(_p(np.ascontiguousarray(y.reshape(-1)), _f32p), _p(fast, _f32p), _p(rate, _f32p), C.byref(r), 4096,
                                _p(cnt, _f32p), _p(read, _f32p), C.c_float(float(d["row_z"].reshape(-1)[0])), _p(q, _f64p),
                                C.c_float(lo), C.c_float(hi))
            assert fast.tobytes() == np.clip(want.reshape(-1), np.float32(lo), np.float32(hi)).tobytes(), (lo, hi)
        ysc = (y.reshape(-1) * np.float32(p["wp"] - p["bl"])) / np.float32(p["ratio"])
        assert np.allclose(rate, ysc.astype(np.float64) / p["K"], rtol=3e-7)           # invK32 multiply: 2 roundings from the exact rate


def test_philox_blocks_equal_the_oracle_philox(core):
    rs = np.random.RandomState(4)
    n = 4096
    index = rs.randint(0, 2 ** 48, size=n, dtype=np.uint64)
    seed, offset = 0x1234_5678_9ABC_DEF0, 0x0BAD_CAFE_0000_0007
    for stream, sub in ((0, 0), (0, 1), (0, 2), (1, 0)):
        out = np.empty((n, 4), np.uint32)
        core.emul_philox_blocks(C.c_uint64(seed), C.c_uint64(offset), _p(index, _u64p), stream, sub, n, _p(out, _u32p))
        ctr = np.stack([index & 0xFFFFFFFF, ((index >> 32) & 0xFFFF) | (sub << 16) | (stream << 24),
                        np.full(n, offset & 0xFFFFFFFF, np.uint64), np.full(n, offset >> 32, np.uint64)], axis=1).astype(np.uint32)
        want = O.philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, seed >> 32], np.uint32))
        assert np.array_equal(out, want)


def _poisson_table():
    t = np.zeros((161, 37), np.float32)
    for r in range(161):
        t[r, 5:] = np.minimum(stats.poisson.cdf(np.arange(32), r / 16.0), 1.0).astype(np.float32)
    return t


def test_sampler_sources_follow_their_distributions(core):
    rs = np.random.RandomState(6)
    n = 1_000_000
    w = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    z = np.empty(n, np.float32)
    core.emul_normal_icdf(_p(w, _u32p), n, _p(z, _f32p))
    assert stats.kstest(z[:200_000], "norm").pvalue > 1e-3
    u = (w.astype(np.float64) + 0.5) * 2.0 ** -32
    ref = np.where(u < 0.5, stats.norm.ppf(u), -stats.norm.ppf(1 - u))
    assert np.abs(z - ref).max() < 3e-6                                            # the inversion itself, word by word
    for lam in (0.15, -0.2, 0.0, 1.0):
        core.emul_tukey_lambda(_p(w, _u32p), C.c_float(lam), n, _p(z, _f32p))
        cell = ((w >> 12).astype(np.float64) + 0.5) * 2.0 ** -20                   # body cells invert their centre
        tail = ((w >> 12) < 256) | ((w >> 12) >= 2 ** 20 - 256)
        uu = np.where(tail, (w.astype(np.float64) + 0.5) * 2.0 ** -32, cell)
        want = stats.tukeylambda.ppf(uu, lam)
        assert np.abs(z - want)[~tail].max() < 2e-5 * max(1.0, np.abs(want[~tail]).max())
        assert np.allclose(z[tail], want[tail], rtol=2e-4, atol=1e-5)
        assert stats.kstest(z[:200_000], lambda x: stats.tukeylambda.cdf(x, lam)).pvalue > 1e-3
    T = _poisson_table()
    out = np.empty(n, np.float32)
    for lam in (0.04, 0.9, 4.37, 9.99, 10.0, 17.3, 64.0, 700.0):
        core.emul_poisson(_p(np.full(n, lam, np.float32), _f32p), _p(w, _u32p), _p(T, _f32p), n, _p(out, _f32p))
        k = out.astype(np.int64)
        assert (k == out).all() and k.min() >= 0
        lo, hi = int(stats.poisson.ppf(1e-6, lam)), int(stats.poisson.ppf(1 - 1e-6, lam)) + 1
        obs = np.bincount(np.clip(k, lo, hi) - lo, minlength=hi - lo + 1).astype(np.float64)
        pm = stats.poisson.pmf(np.arange(lo, hi + 1), lam)
        pm[0] += stats.poisson.cdf(lo - 1, lam)
        pm[-1] += stats.poisson.sf(hi, lam)
        keep = pm * n > 5
        chi2 = ((obs[keep] - pm[keep] * n) ** 2 / (pm[keep] * n)).sum()
        assert chi2 < stats.chi2.ppf(1 - 1e-4, keep.sum() - 1), (lam, chi2)
    lam_mix = rs.uniform(0, 40, n).astype(np.float32)                              # per-element rates, both samplers
    core.emul_poisson(_p(lam_mix, _f32p), _p(w, _u32p), _p(T, _f32p), n, _p(out, _f32p))
    resid = (out - lam_mix) / np.sqrt(np.maximum(lam_mix, 1e-3))
    assert abs(resid[lam_mix > 0.5].mean()) < 5e-3 and abs(resid[lam_mix > 0.5].var() - 1) < 1e-2
    q64, q32 = np.empty(n), np.empty(n, np.float32)
    core.emul_quant_draws(_p(w, _u32p), n, _p(q64, _f64p), _p(q32, _f32p))
    assert np.array_equal(q64, ((w & 0xFFF).astype(np.float64) + 0.5) / 4096 - 0.5) and np.array_equal(q32, (w & 0xFFF).astype(np.float32) / 4096)


def test_pack_and_unpack_sources_are_bit_exact_vs_reference_goldens(core, golden, meta):
    """pack.cu's per-sample arithmetic (csrc/pack_core.cuh) over EVERY sensor code: equals the tables the unmodified reference
    produced, with per-plane bias, without normalisation, and bayer2raw(raw2bayer(x)) == clip(x, bl, wp)."""
    import hashlib
    g = golden("pack")
    _u16p = C.POINTER(C.c_uint16)

    def norm(v, black, wp, do_norm, clip):
        v = np.ascontiguousarray(v, np.float32)
        out = np.empty_like(v)
        core.emul_norm_one(_p(v, _f32p), v.size, C.c_double(black), C.c_double(wp), int(do_norm), int(clip), _p(out, _f32p))
        return out

    def quant(v, wp, bl):
        v = np.ascontiguousarray(v, np.float32)
        out = np.empty(v.shape, np.uint16)
        core.emul_quant_one(_p(v, _f32p), v.size, C.c_float(wp - bl), C.c_float(bl), out.ctypes.data_as(_u16p))
        return out

    for cam, wp, bl, n in (("sony", 16383, 512, 16384), ("imx686", 1023, 64, 1024)):
        codes = np.arange(n, dtype=np.float32)
        for clip in (0, 1):
            t = norm(codes, bl, wp, True, clip)
            assert np.array_equal(t, g[f"{cam}_clip{clip}"])
            assert hashlib.sha1(t.tobytes()).hexdigest()[:16] == meta[f"pack_sha1_{cam}_clip{clip}"]
        rt = quant(norm(codes, bl, wp, True, True), wp, bl)
        assert np.array_equal(rt, np.clip(np.arange(n), bl, wp).astype(np.uint16))
    raw = g["rand_raw"]
    planes = lambda a: np.stack([a[0::2, 0::2], a[0::2, 1::2], a[1::2, 1::2], a[1::2, 0::2]])     # R G1 B G2 (isp_ops.py:87-90)
    got = np.stack([norm(planes(raw)[c], 512 + b, 16383, True, True) for c, b in enumerate((1, -2, 3, 0))])
    assert got.tobytes() == g["rand_packed_bias"].tobytes()
    got = np.stack([norm(planes(raw)[c], 512, 16383, False, False) for c in range(4)])
    assert got.tobytes() == g["rand_packed_nonorm"].tobytes()
    u = g["unpack_in"].reshape(4, -1) if g["unpack_in"].ndim == 3 else g["unpack_in"].reshape(-1, 4, *g["unpack_in"].shape[-2:])[0].reshape(4, -1)
    h, w = g["unpack_in"].shape[-2:]
    q = quant(u, 16383, 512).reshape(4, h, w)
    mosaic = np.empty((2 * h, 2 * w), np.uint16)
    mosaic[0::2, 0::2], mosaic[0::2, 1::2], mosaic[1::2, 1::2], mosaic[1::2, 0::2] = q
    assert np.array_equal(mosaic, g["unpack_out"].reshape(2 * h, 2 * w))
