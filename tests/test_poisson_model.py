"""CPU checks of the two approximations inside the Philox sampler (no GPU needed):
  * normal_icdf: the float32 polynomial model vs scipy.special.ndtri;
  * Poisson lam >= 10: k = floor(Q_N3(Phi^-1(u))) — exact CDF of this sampler vs the exact Poisson CDF.
These are the error bounds quoted in csrc/noise_core.cuh and DESIGN.md."""
import os
import sys

import numpy as np
from scipy import special as sc

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import fit_normal_icdf as F   # noqa: E402

# coefficients as written in csrc/noise_core.cuh
PC = [8.862264752e-01, 1.608277857e-01, 5.522758700e-03, -7.453467697e-04, -4.943624299e-05, 1.418562169e-05,
      -1.025935489e-06, 2.674857846e-08]
PT = [1.426639557e-01, 5.167053342e-01, 1.309477687e-01, -2.902236022e-02, 3.719373606e-03, -2.602138266e-04,
      7.718497727e-06]


def test_coefficients_in_the_kernel_source_match():
    src = open(os.path.join(ROOT, "pnnp_b200", "csrc", "noise_core.cuh")).read()
    for c in PC + PT:
        assert f"{abs(c):.9e}f" in src, c


def test_normal_icdf_float32_model():
    rs = np.random.RandomState(1)
    words = np.concatenate([rs.randint(0, 2 ** 32, size=1_000_000, dtype=np.uint64), np.arange(0, 2048, dtype=np.uint64),
                            (2 ** 32 - 1 - np.arange(0, 2048)).astype(np.uint64)])
    z = F.model_f32(words, np.array(PC), np.array(PT)).astype(np.float64)
    u = (words.astype(np.float64) + 0.5) * 2.0 ** -32
    ref = np.where(u < 0.5, sc.ndtri(u), -sc.ndtri(1 - u))
    assert np.abs(z - ref).max() < 2e-6
    assert np.isfinite(z).all() and z.min() < -6.1 and z.max() > 6.1


def _qn3(w, lam):
    s = np.sqrt(lam)
    return lam + s * w + (1 / 3 + w * w / 6) + (-w / 36 - w ** 3 / 72) / s + (-8 / 405 + 7 * w * w / 810 + w ** 4 / 270) / lam


def _dqn3(w, lam):
    s = np.sqrt(lam)
    return s + w / 3 + (-1 / 36 - w * w / 24) / s + (14 * w / 810 + 4 * w ** 3 / 270) / lam


def sampler_cdf_error(lam):
    """max_k |P_sampler(X <= k-1) - P_poisson(X <= k-1)|: the sampler returns <= k-1 iff Q(z) < k."""
    k = np.arange(max(1, int(lam - 9 * np.sqrt(lam))), int(lam + 10 * np.sqrt(lam) + 12)).astype(float)
    q, p = sc.gammaincc(k, lam), sc.gammainc(k, lam)                 # exact P(X <= k-1) and its complement
    keep = (q > 1e-13) & (p > 1e-13)
    k, q, p = k[keep], q[keep], p[keep]
    w_exact = np.where(q < 0.5, sc.ndtri(q), -sc.ndtri(p))
    w = w_exact.copy()
    for _ in range(40):                                               # solve Q(w) = k
        w = w - (_qn3(w, lam) - k) / _dqn3(w, lam)
    return np.abs(sc.ndtr(w) - sc.ndtr(w_exact)).max(), np.abs(sc.ndtr(w) - sc.ndtr(w_exact)).sum()


def test_poisson_inversion_cdf_error_bounds():
    worst = {}
    for lam in (10.0, 10.5, 12.0, 16.0, 24.0, 32.0, 64.0, 128.0, 400.0, 835.0):
        worst[lam] = sampler_cdf_error(lam)
    assert worst[10.0][0] < 5e-5 and worst[10.0][1] < 5e-4        # KS distance, total-variation bound at the switch point
    assert worst[32.0][0] < 5e-6 and worst[64.0][0] < 1e-6 and worst[835.0][0] < 1e-8
    assert all(worst[a][0] >= worst[b][0] for a, b in zip(sorted(worst)[:-1], sorted(worst)[1:]))   # monotone in lam


# ---------------------------------------------------------------------------------------------------------------
# lam < 10: the table form of the exact inversion (csrc/noise_core.cuh: poisson_small_table), modelled in float32
# ---------------------------------------------------------------------------------------------------------------
ROWS, COLS, PAD = 161, 32, 5


def _table():
    from scipy import stats
    t = np.zeros((ROWS, PAD + COLS), dtype=np.float32)
    for r in range(ROWS):
        t[r, PAD:] = np.minimum(stats.poisson.cdf(np.arange(COLS), r / 16.0), 1.0).astype(np.float32)
    return t


def _table_sampler(lam, words, t):
    f32 = np.float32
    lam = lam.astype(f32)
    u = (((words >> np.uint64(9)).astype(np.float64) + 0.5) * 2.0 ** -23).astype(f32)                  # top 23 bits, exact in float32
    i = (lam * f32(16)).astype(np.int64)
    delta = (lam.astype(np.float64) - 0.0625 * i).astype(f32)                                          # exact
    k = np.zeros(lam.shape, dtype=np.int64)
    rows = t[i]
    for s in (16, 8, 4, 2, 1):
        k += np.where(rows[np.arange(len(k)), PAD + k + s - 1] < u, s, 0)
    c2 = f32(0.5) * delta * delta
    c3 = c2 * delta * f32(0.33333334)
    c4 = c3 * delta * f32(0.25)
    ed = np.exp(delta.astype(np.float64)).astype(f32)
    ue = u * ed
    done = np.zeros(lam.shape, dtype=bool)
    for _ in range(COLS):
        kk = np.minimum(k, COLS - 1)
        g = lambda j: rows[np.arange(len(k)), PAD + kk - j].astype(np.float64)
        s_ = (g(0) + delta * g(1) + c2 * g(2) + c3 * g(3) + c4 * g(4)).astype(f32)
        ok = (s_ >= ue) | (k >= COLS)
        done |= ok
        k = np.where(done, k, k + 1)
        if done.all():
            break
    return k


def test_kernel_source_uses_the_modelled_table_geometry():
    src = open(os.path.join(ROOT, "pnnp_b200", "csrc", "noise_core.cuh")).read()
    assert "kPoisRows = 161, kPoisCols = 32, kPoisPad = 5" in src and "lam * 16.0f" in src


def test_small_rate_table_inversion_is_the_exact_inverse_cdf():
    """k(u, lam) from the table algorithm == scipy's exact Poisson quantile, except within float32 rounding of a CDF
    step (|F(k) - u| < 4e-7), for rates across [0, 10) including grid points and the ends of the grid cells."""
    from scipy import stats
    t = _table()
    rs = np.random.RandomState(3)
    n = 400_000
    lam = np.concatenate([rs.uniform(0, 10, n), rs.randint(0, 160, 20_000) / 16.0, rs.randint(1, 160, 20_000) / 16.0 - 1e-6,
                          rs.uniform(0, 0.05, 20_000)]).astype(np.float32)
    lam = np.clip(lam, 0, np.float32(9.999999))
    words = rs.randint(0, 2 ** 32, size=lam.size, dtype=np.uint64)
    k = _table_sampler(lam, words, t)
    u = np.minimum(words.astype(np.float64) * 2.0 ** -32 + 2.0 ** -33, 0.99999994)
    ref = stats.poisson.ppf(u, lam.astype(np.float64)).astype(np.int64)
    bad = np.nonzero(k != ref)[0]
    assert bad.size < 2e-5 * lam.size, bad.size                                  # a handful of boundary cases
    if bad.size:
        lo, hi = np.minimum(k[bad], ref[bad]), np.maximum(k[bad], ref[bad])
        assert (hi - lo == 1).all()
        assert np.abs(stats.poisson.cdf(lo, lam[bad].astype(np.float64)) - u[bad]).max() < 4e-7
    assert k.max() < COLS                                                        # the sequential fallback was not needed here
