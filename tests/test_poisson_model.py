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
