"""Fused noise synthesis on the GPU through the C ABI.

  * replay mode + the reference's own draws  → bit-exact vs the reference goldens and the oracle;
  * Philox mode                               → same arithmetic core (debug draws replayed give the
    same bits), Philox words equal the CPU Philox, per-stage KS tests / moments against the exact
    distributions the reference samples from, shard-independence.
"""
import numpy as np
import pytest
import torch
from scipy import stats

import oracle_np as O
import pnnp_b200 as P
from pnnp_b200 import _lib
from conftest import decode_param

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _np_draws(g, tag):
    d = {}
    if f"{tag}_counts" in g.files:
        d["shot"] = g[f"{tag}_counts"]
    if f"{tag}_shot_z" in g.files:
        d["shot"] = g[f"{tag}_shot_z"]
    for k in ("read", "row_z", "q"):
        if f"{tag}_{k}" in g.files:
            d[k] = g[f"{tag}_{k}"]
    return d


def test_replay_bit_exact_vs_reference_goldens_numpy_chain(golden, meta):
    g = golden("noisy_obs")
    y = _cuda(g["y"])[None]
    for c in meta["noisy_obs_cases"]:
        p = decode_param(c["param"])
        d = _np_draws(g, c["tag"])
        d = {k: (v[None] if k != "row_z" else v[None]) for k, v in d.items()}
        out = P.replay_batch(y, [p], c["code"], d, chain=_lib.CHAIN_NUMPY, ori=c["ori"], clip=c["clip"])
        assert out[0].cpu().numpy().tobytes() == g[c["tag"] + "_z"].tobytes(), c


def test_replay_bit_exact_vs_reference_goldens_torch_chain(golden, meta):
    g = golden("noisy_torch")
    y = _cuda(g["y"])[None]
    for c in meta["noisy_torch_cases"]:
        p = decode_param(c["param"])
        t = c["tag"]
        d = {"shot": g[t + "_counts"][None], "read": g[t + "_read"][None]}
        if t + "_row_z" in g.files:
            d["row_z"] = g[t + "_row_z"][None]
        if t + "_q_u" in g.files:
            d["q"] = g[t + "_q_u"][None].astype(np.float64)
        out = P.replay_batch(y, [p], c["code"], d, chain=_lib.CHAIN_TORCH, ori=c["ori"], clip=bool(c["clip"]))
        assert out[0].cpu().numpy().tobytes() == g[t + "_z"].tobytes(), c


@pytest.mark.parametrize("code", ["pgrq", "prq", "grq", "pg"])
def test_replay_bit_exact_vs_oracle_crop_size(code):
    """Reference-order NumPy draws on 3 crops of 4x512x512 with different parameter regimes."""
    rs = np.random.RandomState(11)
    y = (rs.rand(3, 4, 512, 512).astype(np.float32)) ** 2
    np.random.seed(21)
    params = [O.sample_params("SonyA7S2"), O.sample_params_max("SonyA7S2", iso=6400), O.sample_params_max("IMX686", iso=100)]
    want, draws = [], {"shot": [], "read": [], "row_z": [], "q": []}
    for i, p in enumerate(params):
        np.random.seed(50 + i)
        z, d = O.generate_noisy_obs(y[i], param=p, noise_code=code, return_draws=True)
        want.append(z)
        draws["shot"].append(d["counts"] if "counts" in d else d["shot_z"])
        draws["read"].append(d["read"])
        draws["row_z"].append(d.get("row_z", np.zeros((4, 512, 1), np.float32)))
        draws["q"].append(d.get("q", np.zeros((4, 512, 512))))
    draws = {k: np.stack(v) for k, v in draws.items()}
    out = P.replay_batch(_cuda(y), params, code, draws, chain=_lib.CHAIN_NUMPY).cpu().numpy()
    assert out.tobytes() == np.stack(want).tobytes()


def _mk(n, h, w, seed, dark=True):
    rs = np.random.RandomState(seed)
    y = rs.rand(n, 4, h, w).astype(np.float32)
    return y ** 2 if dark else y


@pytest.mark.parametrize("chain,code", [(_lib.CHAIN_NUMPY, "pgrq"), (_lib.CHAIN_NUMPY, "gq"), (_lib.CHAIN_NUMPY, "pgrqd"),
                                        (_lib.CHAIN_NUMPY, "p"), (_lib.CHAIN_TORCH, "prq"), (_lib.CHAIN_TORCH, "pb")])
@pytest.mark.parametrize("w", [512, 100, 37])
def test_philox_kernel_uses_the_replay_arithmetic(chain, code, w):
    """synth(debug) draws → replay → identical bits; proves the Philox kernel and the bit-exact
    replay kernel share one arithmetic core (also covers the scalar, non-multiple-of-4 path)."""
    n, h = 3, 24
    y = _cuda(_mk(n, h, w, 3))
    np.random.seed(4)
    if chain == _lib.CHAIN_NUMPY:
        params = [O.sample_params("SonyA7S2"), O.sample_params_max("SonyA7S2", iso=3200), O.sample_params_max("IMX686", iso=6400)]
        if "d" in code:
            params[1]["bias"] = np.array([0.5, -0.25, 1.0, 2.0])
    else:
        params = [O.sample_params_max("SonyA7S2") for _ in range(n)]
    gen = P.PhiloxGenerator(1234)
    out, d = P.synthesize_batch(y, params, code, chain, generator=gen, debug=True)
    gen2 = P.PhiloxGenerator(1234)
    out2 = P.synthesize_batch(y, params, code, chain, generator=gen2)
    assert torch.equal(out, out2)                                   # debug and production kernels agree
    rep = P.replay_batch(y, params, code, {"shot": d["shot"], "read": d["read"], "row_z": d["row_z"], "q": d["q"]}, chain=chain)
    assert torch.equal(out, rep)
    out3 = P.synthesize_batch(y, params, code, chain, generator=gen2)  # offset advanced → fresh draws
    assert not torch.equal(out, out3)


def test_specialised_kernel_is_bit_identical_to_replay_at_scale():
    """16 crops x 4x512x512 (16.8 M elements) through the branch-free instantiation (reciprocal-multiply
    divisions with Markstein correction) == the generic replay kernel (IEEE divisions) on the same draws."""
    g = torch.Generator(device="cuda").manual_seed(3)
    y = torch.rand((16, 4, 512, 512), device="cuda", generator=g) ** 2
    np.random.seed(12)
    params = [P.sample_params("SonyA7S2") for _ in range(16)]
    out, d = P.synthesize_batch(y, params, "pgrq", generator=P.PhiloxGenerator(5), debug=True)
    rep = P.replay_batch(y, params, "pgrq", {"shot": d["shot"], "read": d["read"], "row_z": d["row_z"], "q": d["q"]})
    assert torch.equal(out, rep)
    # and the generic kernel (hint withheld) agrees with its own replay too
    tab = P.ParamTable(params, y.device)
    tab.uniform_f64 = False
    out2, d2 = P.synthesize_batch(y[:2].contiguous(), None, "pgrq", generator=P.PhiloxGenerator(5), debug=True, table=tab)
    rep2 = P.replay_batch(y[:2].contiguous(), params[:2], "pgrq", {"shot": d2["shot"], "read": d2["read"], "row_z": d2["row_z"], "q": d2["q"]})
    assert torch.equal(out2, rep2)


def test_device_philox_words_match_cpu_philox():
    """Draw layout (csrc/noise_core.cuh): group G = g >> 2 owns Philox blocks sub 0 (shot words) and sub 1 ("mix" words:
    read-noise cell = mix >> 12, quantisation draw = mix & 0xFFF).  The quantisation draw exposes the low 12 bits of word
    e = g & 3 of block sub 1 exactly: q = (bits + 0.5) / 4096 - 0.5."""
    n, c, h, w = 2, 4, 8, 64
    y = _cuda(_mk(n, h, w, 5))
    np.random.seed(1)
    params = [O.sample_params("SonyA7S2") for _ in range(n)]
    seed, crop0 = 0x0123456789ABCDEF, 5
    gen = P.PhiloxGenerator(seed)
    gen.offset = 7
    _, d = P.synthesize_batch(y, params, "pgrq", generator=gen, crop_id0=crop0, debug=True)
    q = d["q"].cpu().numpy().reshape(-1)
    bits = (q + 0.5) * 4096.0 - 0.5
    assert np.array_equal(bits, np.round(bits)) and bits.min() >= 0 and bits.max() <= 4095
    g = np.arange(n * c * h * w, dtype=np.uint64) + np.uint64(crop0 * c * h * w)
    grp, e = g >> np.uint64(2), (g & np.uint64(3)).astype(np.int64)
    sub = np.ones_like(grp)
    ctr = np.stack([grp & np.uint64(0xFFFFFFFF), ((grp >> np.uint64(32)) & np.uint64(0xFFFF)) | (sub << np.uint64(16)),
                    np.full_like(grp, 7), np.zeros_like(grp)], -1).astype(np.uint32)
    ref = O.philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32))
    mix = ref[np.arange(g.size), e].astype(np.uint64)
    assert np.array_equal(bits.astype(np.uint64), mix & np.uint64(0xFFF))
    # the generic kernel (hint withheld) uses the same draws as the specialised one
    tab = P.ParamTable(params, y.device)
    tab.uniform_f64 = False
    gen2 = P.PhiloxGenerator(seed)
    gen2.offset = 7
    out_f = P.synthesize_batch(y, params, "pgrq", generator=P.PhiloxGenerator(seed), crop_id0=crop0)
    out_g = P.synthesize_batch(y, None, "pgrq", generator=P.PhiloxGenerator(seed), crop_id0=crop0, table=tab)
    assert torch.equal(out_f, out_g)


def test_read_noise_tail_refinement_keeps_full_resolution():
    """The outer 256 cells of each tail of the 20-bit read-noise draw are refined with 12 more random bits: extreme
    quantiles are not confined to the cell-centre lattice (2 x 2^21 draws -> ~2000 tail draws, almost all distinct)."""
    y = torch.zeros((1, 4, 1024, 1024), device="cuda")
    _, d = P.synthesize_batch(y, [_flat_param(1.0, sigTL=1.0, lam=-0.026)], "pg", generator=P.PhiloxGenerator(77), debug=True)
    r = d["read"].flatten()
    q_lo = float(O.tukeylambda_ppf(np.float64(256.0 / 2 ** 20), -0.026))
    tail = r[r < q_lo]
    assert 600 < tail.numel() < 1500                               # expected 4 Mi * 2^-12 = 1024
    assert torch.unique(tail).numel() > 0.95 * tail.numel()        # cell centres alone would give <= 256 distinct values


def test_shard_independence():
    """Crops [0,6) in one launch == crops [0,2) and [2,6) launched separately (any GPU count)."""
    y = _cuda(_mk(6, 32, 64, 8))
    np.random.seed(2)
    params = [O.sample_params("SonyA7S2") for _ in range(6)]
    full = P.synthesize_batch(y, params, "pgrq", generator=P.PhiloxGenerator(99))
    a = P.synthesize_batch(y[:2].contiguous(), params[:2], "pgrq", generator=P.PhiloxGenerator(99), crop_id0=0)
    b = P.synthesize_batch(y[2:].contiguous(), params[2:], "pgrq", generator=P.PhiloxGenerator(99), crop_id0=2)
    assert torch.equal(full, torch.cat([a, b]))


def _flat_param(K, sigTL=1.0, lam=-0.026, sigR=0.5, sigGs=2.0, ratio=1.0, wp=16383, bl=512):
    return {"K": np.float64(K), "sigTL": np.float64(sigTL), "sigR": np.float64(sigR), "sigGs": np.float64(sigGs),
            "bias": np.float64(0.0), "lam": lam, "q": 1 / 2 ** 14, "ratio": ratio, "wp": wp, "bl": bl}


@pytest.mark.parametrize("lam", [0.05, 0.7, 3.0, 9.5, 10.5, 25.0, 99.0, 400.0, 830.0])
def test_poisson_stage_ks_and_moments(lam):
    """Poisson counts for a constant rate vs the exact Poisson CDF (discrete KS, n = 2^21)."""
    n_el = 4 * 512 * 1024
    K = 1.0
    yval = np.float32(lam * K / 15871.0)
    y = torch.full((1, 4, 512, 1024), float(yval), device="cuda")
    _, d = P.synthesize_batch(y, [_flat_param(K)], "p", generator=P.PhiloxGenerator(int(lam * 100)), debug=True)
    k = d["shot"].cpu().numpy().reshape(-1).astype(np.int64)
    lam_eff = float(np.float32(np.float32(yval * np.float32(15871.0)) / np.float32(1.0)) / K)
    assert k.min() >= 0
    kmax = int(k.max())
    emp = np.cumsum(np.bincount(k, minlength=kmax + 1)) / n_el
    cdf = stats.poisson.cdf(np.arange(kmax + 1), lam_eff)
    D = np.abs(emp - cdf).max()
    assert D < 1.63 / np.sqrt(n_el), (lam, D)                    # alpha = 0.01 (conservative for discrete)
    assert abs(k.mean() - lam_eff) < 5 * np.sqrt(lam_eff / n_el)
    assert abs(k.var() / lam_eff - 1) < 5 * np.sqrt(2.0 / n_el + 1.0 / (lam_eff * n_el)) + 1e-3


def test_poisson_rate_follows_pixels():
    """Per-pixel rates (dark-scene distribution): E[count] = lam, Var = lam, over many crops."""
    rs = np.random.RandomState(0)
    yv = (rs.rand(1, 4, 64, 64).astype(np.float32) ** 2)
    y = _cuda(np.repeat(yv, 256, axis=0))
    p = _flat_param(2.5, ratio=100.0)
    _, d = P.synthesize_batch(y, [p] * 256, "p", generator=P.PhiloxGenerator(5), debug=True)
    k = d["shot"].cpu().numpy().astype(np.float64)
    lam = O.lam_of(yv[0], p)[1]
    z = (k.mean(0) - lam) / np.sqrt(np.maximum(lam, 1e-9) / 256)
    assert abs(z[lam > 0.5].mean()) < 0.05 and abs(z[lam > 0.5].std() - 1) < 0.05
    v = k.var(0, ddof=1)
    assert abs((v[lam > 1] / lam[lam > 1]).mean() - 1) < 0.01


@pytest.mark.parametrize("lam_tl", [-0.26, -0.026, 0.0005, 0.102, 0.1474653])
def test_tukey_lambda_stage_ks(lam_tl):
    y = torch.zeros((1, 4, 512, 512), device="cuda")
    sig = 1.7
    _, d = P.synthesize_batch(y, [_flat_param(1.0, sigTL=sig, lam=lam_tl)], "pg", generator=P.PhiloxGenerator(3), debug=True)
    r = d["read"].cpu().numpy().reshape(-1).astype(np.float64)
    D, pval = stats.kstest(r / sig, lambda x: stats.tukeylambda.cdf(x, lam_tl))
    assert pval > 1e-3, (lam_tl, D, pval)
    ref = stats.tukeylambda.rvs(lam_tl, scale=sig, size=r.size, random_state=np.random.RandomState(1))
    assert abs(np.std(r) / np.std(ref) - 1) < 0.02
    assert abs(np.quantile(r, 0.999) / np.quantile(ref, 0.999) - 1) < 0.05


def test_gaussian_read_row_and_quant_stage_ks():
    y = torch.zeros((8, 4, 256, 256), device="cuda")
    p = _flat_param(1.0, sigGs=3.0, sigR=0.5)
    _, d = P.synthesize_batch(y, [p] * 8, "prq", generator=P.PhiloxGenerator(8), debug=True)
    read = d["read"].cpu().numpy().reshape(-1).astype(np.float64)[:1 << 20]
    assert stats.kstest(read / 3.0, "norm").pvalue > 1e-3
    rowz = d["row_z"].cpu().numpy().reshape(-1).astype(np.float64)
    assert rowz.size == 8 * 4 * 256 and stats.kstest(rowz, "norm").pvalue > 1e-3
    q = d["q"].cpu().numpy().reshape(-1)[:1 << 20]
    assert stats.kstest(q + 0.5, "uniform").pvalue > 1e-3 and q.min() >= -0.5 and q.max() <= 0.5
    # Gaussian shot approximation (no 'p'): standard normal draw
    _, d = P.synthesize_batch(y[:1], [p], "g", generator=P.PhiloxGenerator(9), debug=True)
    assert stats.kstest(d["shot"].cpu().numpy().reshape(-1).astype(np.float64), "norm").pvalue > 1e-3
    # independence of the four streams of one element block
    c = np.corrcoef(np.stack([read, q]))
    assert np.abs(c - np.eye(2)).max() < 0.01


def test_row_noise_is_constant_along_width_and_differs_across_rows():
    y = torch.zeros((2, 4, 64, 2128), device="cuda")      # 5 segments per row: same draw in each
    p = _flat_param(1.0, sigR=4.0, sigGs=0.0)       # Poisson(0) = 0 and N(0, 0) = 0: only the row term is left
    out = P.synthesize_batch(y, [p, p], "pr", generator=P.PhiloxGenerator(2), ori=True)
    o = out.cpu().numpy()
    assert np.all(o == o[..., :1]) and len(np.unique(o[..., 0])) == 2 * 4 * 64


def test_end_to_end_matches_reference_statistics():
    """Whole-function two-sample KS: product generate_noisy_obs vs the oracle's reference-order
    sampling.  The pooled-residual KS uses 'pgq' (i.i.d. per pixel); with 'r' every pixel of a row
    shares one draw, so pooled residuals are not independent and only moments are compared."""
    rs = np.random.RandomState(4)
    y = (rs.rand(4, 256, 256).astype(np.float32) ** 2)
    np.random.seed(9)
    p = O.sample_params("SonyA7S2")
    for code, do_ks in (("pgq", True), ("pgrq", False)):
        np.random.seed(10)
        want = O.generate_noisy_obs(y, param=p, noise_code=code)
        P.manual_seed(77)
        got = P.generate_noisy_obs(y, param=p, noise_code=code)
        assert got.dtype == np.float32 and got.shape == want.shape
        res_w, res_g = (want - y).reshape(-1), (got - y).reshape(-1)
        if do_ks:
            assert stats.ks_2samp(res_w, res_g).pvalue > 1e-3
            assert abs(res_g.mean() - res_w.mean()) < 8 * res_w.std() / np.sqrt(res_w.size)
        assert abs(res_g.std() / res_w.std() - 1) < 0.02
    lo = -p["bl"] / p["wp"] * p["ratio"]
    assert got.min() >= np.float32(lo) - 1e-6


def test_full_config2_size_properties():
    """BASELINE config 2 at full size (64 x 4 x 512 x 512): bounds, determinism, fused post-clip."""
    g = torch.Generator(device="cuda").manual_seed(1997)
    y = torch.rand((64, 4, 512, 512), device="cuda", generator=g) ** 2
    np.random.seed(1997)
    params = [P.sample_params("SonyA7S2") for _ in range(64)]
    a = P.synthesize_batch(y, params, "pgrq", generator=P.PhiloxGenerator(1), post_clip=(-np.inf, 1.0))
    b = P.synthesize_batch(y, params, "pgrq", generator=P.PhiloxGenerator(1))
    assert torch.equal(a, b.clamp(max=1.0))                       # fused follow-up clamp == separate clamp
    ratios = torch.tensor([p["ratio"] for p in params], device="cuda").view(-1, 1, 1, 1)
    assert bool((b >= (-512 / 16383) * ratios - 1e-4).all()) and bool((b <= ratios + 1e-3).all())
    assert abs(float((b - y).mean())) < 2e-3                       # noise is zero-mean up to clipping
    c = P.synthesize_batch(y, params, "pgrq", generator=P.PhiloxGenerator(1))
    assert torch.equal(b, c)


def test_reference_signatures_roundtrip():
    rs = np.random.RandomState(1)
    y = rs.rand(4, 64, 64).astype(np.float32)
    np.random.seed(3)
    p = P.sample_params_max("SonyA7S2")
    tp = {k: torch.from_numpy(np.array(v, np.float32)).cuda() for k, v in p.items()}
    yt = torch.from_numpy(y).cuda()
    z = P.generate_noisy_torch(yt, param=tp, noise_code="prq", ori=False, clip=2)
    assert z.shape == yt.shape and z.is_cuda and float(z.min()) >= 0.0
    z2 = P.generate_noisy_obs(yt, param=p, noise_code="pgr")
    assert z2.is_cuda and z2.dtype == torch.float32


def test_tiny_tukey_lambda_withholds_the_specialised_kernel_hint():
    """The specialised kernel has only the power form of the Tukey-lambda quantile; PNNP_CODE_UNIFORM_F64 therefore also promises
    |lam| >= 1e-3 (include/pnnp_b200.h).  A table with a smaller shape parameter goes through the generic kernel (series form) and
    still equals the replay of its own draws; the read noise is the logistic-like limit, finite and centred."""
    g = torch.Generator(device="cuda").manual_seed(9)
    y = torch.rand((2, 4, 64, 512), device="cuda", generator=g) ** 2
    np.random.seed(4)
    params = [P.sample_params("SonyA7S2") for _ in range(2)]
    assert P.ParamTable(params, y.device).uniform_f64
    params[1]["lam"] = 5e-4
    tab = P.ParamTable(params, y.device)
    assert not tab.uniform_f64
    out, d = P.synthesize_batch(y, None, "pgrq", generator=P.PhiloxGenerator(5), debug=True, table=tab)
    rep = P.replay_batch(y, params, "pgrq", {"shot": d["shot"], "read": d["read"], "row_z": d["row_z"], "q": d["q"]})
    assert torch.equal(out, rep)
    r = d["read"][1].double()
    assert torch.isfinite(r).all() and abs(float(r.mean())) < 0.05 * float(r.std())
