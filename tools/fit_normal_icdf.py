"""Derives the polynomial coefficients of pnnp::normal_icdf (csrc/noise_core.cuh): Chebyshev-node
least-squares fits of erfinv in the variable t = -log2(4 p (1-p)) (central) and s = sqrt(t) (tail),
and checks the float32 evaluation against scipy.special.ndtri.  Run: python tools/fit_normal_icdf.py"""
import numpy as np
from numpy.polynomial import chebyshev as Ch, polynomial as Po
from scipy import special as sc

T0, TMAX = 8.25, 34.0


def fit(fun, a, b, deg, n=4000):
    k = np.arange(n)
    z = np.cos(np.pi * (k + 0.5) / n)
    t = 0.5 * (b - a) * z + 0.5 * (b + a)
    pz = Ch.cheb2poly(Ch.chebfit(z, fun(t), deg))
    lin = np.array([-(a + b) / (b - a), 2 / (b - a)])
    pt = np.array([0.0])
    for i, ci in enumerate(pz):
        term = np.array([1.0])
        for _ in range(i):
            term = Po.polymul(term, lin)
        pt = Po.polyadd(pt, ci * term)
    return pt


def g_central(t):
    x = np.sqrt(np.maximum(1 - 2.0 ** (-t), 1e-300))
    r = sc.erfinv(x) / x
    r[t < 1e-12] = np.sqrt(np.pi) / 2
    return r


def g_tail(s):
    om = 2.0 ** (-(s * s))
    x = np.sqrt(1 - om)
    return sc.erfcinv(om / (1 + x))


def model_f32(words, pc, pt):
    """float32 model of the device function: word (uint32) -> standard normal."""
    f = np.float32
    w = words.astype(np.uint64)
    neg = w < 2 ** 31                                  # lower half -> negative quantile
    tmin = np.where(neg, w, (2 ** 32 - 1) - w)        # min(word, ~word)
    p = (tmin.astype(np.float32) * f(2.0 ** -32) + f(2.0 ** -33)).astype(np.float32)      # (0, 0.5]
    om = (f(4.0) * p * (f(1.0) - p)).astype(np.float32)
    t = (-np.log2(om.astype(np.float64))).astype(np.float32)
    x = (f(1.0) - f(2.0) * p).astype(np.float32)
    cen = np.zeros_like(t)
    for c in pc[::-1]:
        cen = (cen * t + f(c)).astype(np.float32)
    cen = (cen * x).astype(np.float32)
    s = np.sqrt(t).astype(np.float32)
    tl = np.zeros_like(t)
    for c in pt[::-1]:
        tl = (tl * s + f(c)).astype(np.float32)
    e = np.where(t < f(T0), cen, tl)
    z = (f(np.sqrt(2.0)) * e).astype(np.float32)
    return np.where(neg, -z, z)


if __name__ == "__main__":
    pc = fit(g_central, 0.0, T0, 7)
    pt = fit(g_tail, np.sqrt(T0), np.sqrt(TMAX), 6)
    print("central:", ", ".join(f"{np.float32(c):.9e}f" for c in pc))
    print("tail   :", ", ".join(f"{np.float32(c):.9e}f" for c in pt))
    rs = np.random.RandomState(0)
    words = np.concatenate([rs.randint(0, 2 ** 32, size=2_000_000, dtype=np.uint64),
                            np.arange(0, 4096, dtype=np.uint64), 2 ** 32 - 1 - np.arange(0, 4096, dtype=np.uint64),
                            (2 ** 31 - 2048 + np.arange(0, 4096)).astype(np.uint64)]).astype(np.uint64)
    z = model_f32(words, pc, pt).astype(np.float64)
    u = (words.astype(np.float64) + 0.5) * 2.0 ** -32
    ref = np.where(u < 0.5, sc.ndtri(u), -sc.ndtri(1 - u))
    err = np.abs(z - ref)
    print("max abs err", err.max(), "at z =", ref[err.argmax()], "; central max", err[np.abs(ref) < 3].max())
