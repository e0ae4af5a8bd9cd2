"""CPU oracle for the PNNP hot path — TEST INFRASTRUCTURE ONLY.

A NumPy / torch-CPU restatement of the reference algorithms on the hot path
(SURVEY.md §8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file; the product package
``pnnp_b200`` never does and has no CPU fallback.

Pinning.  The reference ships no tests or golden vectors, so the oracle is
pinned against *outputs of the live reference run in the build container*
(``oracle/make_golden.py`` → ``tests/golden/*.npz``; ``tests/test_oracle_vs_reference.py``
re-checks live whenever /root/reference is present).  Versions the goldens were made with
are stored inside each fixture (numpy 2.3.5 / scipy 1.18.1 / torch 2.11.0).

PARITY UNPINNED for one row: ``psnr`` / ``ssim`` (E2) restate scikit-image's documented defaults, but scikit-image is
not installed in the build container, so no output of the reference's own metric calls exists to pin them against
(DESIGN.md §2).  The third-party samplers the reference calls (NumPy ``RandomState.poisson``, SciPy ``tukeylambda.rvs``)
are called here as well, not restated.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
import numpy as np

F32 = np.float32
F64 = np.float64

# --------------------------------------------------------------------------------------
# P1 / P2  Bayer pack / unpack                         utils/isp_ops.py:84-112
# --------------------------------------------------------------------------------------

def raw2bayer(raw, wp=1023, bl=64, norm=True, clip=False, bias=np.array([0, 0, 0, 0])):
    """utils/isp_ops.py:84-96.  Plane order R(0,0) G1(0,1) B(1,1) G2(1,0).  `bias + bl` is an
    int64 array, so subtraction and division run in float64 and round to float32 once."""
    r = np.asarray(raw).astype(F32)
    H, W = r.shape
    planes = [r[0:H:2, 0:W:2], r[0:H:2, 1:W:2], r[1:H:2, 1:W:2], r[1:H:2, 0:W:2]]
    out = np.stack(planes, axis=0).astype(F32)
    if norm:
        black = (np.asarray(bias) + bl).reshape(4, 1, 1)
        out = (out - black) / (wp - black)
    if clip:
        out = np.clip(out, 0, 1)
    return out.astype(F32)


def bayer2raw(packed, wp=16383, bl=512):
    """utils/isp_ops.py:98-112.  clip[0,1] → x*(wp-bl)+bl in float32 → truncating uint16 cast."""
    p = np.asarray(packed, dtype=F32)
    if p.ndim == 4:
        p = p[0]
    p = np.clip(p, 0, 1)
    p = p * (wp - bl) + bl          # python ints are weak: stays float32
    C, h, w = p.shape
    raw = np.empty((2 * h, 2 * w), dtype=np.uint16)
    raw[0::2, 0::2] = p[0]
    raw[0::2, 1::2] = p[1]
    raw[1::2, 1::2] = p[2]
    raw[1::2, 0::2] = p[3]
    return raw


def bayer2rggb(bayer):
    """utils/isp_ops.py:57-59 (HWC, true RGGB order; listed so the two orders are not confused)."""
    H, W = bayer.shape
    return bayer.reshape(H // 2, 2, W // 2, 2).transpose(0, 2, 1, 3).reshape(H // 2, W // 2, 4)


def rggb2bayer(rggb):
    """utils/isp_ops.py:61-63."""
    H, W, _ = rggb.shape
    return rggb.reshape(H, W, 2, 2).transpose(0, 2, 1, 3).reshape(H * 2, W * 2)


# --------------------------------------------------------------------------------------
# S1  camera tables                                    data_process/process.py:215-308
# --------------------------------------------------------------------------------------
# Layout differs from the reference on purpose: one row per camera, column names in _FIT_COLS.
_FIT_COLS = ("Kmin", "Kmax", "lam", "qbits", "wp", "bl",
             "sigTLk", "sigTLb", "sigTLsig", "sigRk", "sigRb", "sigRsig",
             "sigGsk", "sigGsb", "sigGssig",
             "sigReadk", "sigReadb", "sigReadsig", "uReadk", "uReadb", "uReadsig")
_N = None
_FITS = {
    "NikonD850":        (1.2, 2.4828, -0.26, 14, 16383, 512, 0.906, -0.6754, 0.035165, 0.8322, -2.3326, 0.301333,
                         0.8322, -0.1754, 0.035165, _N, _N, _N, _N, _N, _N),
    "IMX686":           (-0.19118, 2.16820, 0.102, 10, 1023, 64, 0.85187, 0.07991, 0.02921, 0.87611, -2.11455, 0.03274,
                         0.85187, 0.67991, 0.02921, _N, _N, _N, _N, _N, _N),
    "SonyA7S2_lowISO":  (-1.67214, 0.42228, -0.026, 14, 16383, 512, 0.74043, 0.86182, 0.00712, 0.78782, -0.34227, 0.02832,
                         0.82966, 1.49343, 0.00359, 0.82879, 1.50601, 0.00362, 0.01472, 0.01129, 0.00034),
    "SonyA7S2_highISO": (0.64567, 2.51606, -0.025, 14, 16383, 512, 0.74901, -0.12348, 0.00638, 0.62945, -1.51040, 0.02609,
                         0.82878, 0.44162, 0.00153, 0.82645, 0.45061, 0.00156, 0.00385, 0.00674, 0.00039),
    "CRVD":             (1.31339, 3.95448, 0.015, 12, 4095, 240, 0.95495, 0.01618, 0.00790, 0.93368, -2.19692, 0.02473,
                         0.95387, 0.01552, 0.00855, _N, _N, _N, _N, _N, _N),
}


def get_camera_noisy_params(camera_type=None):
    """process.py:215-255 — dict of the log-linear fit constants (unknown camera → NikonD850)."""
    row = _FITS.get(camera_type, _FITS["NikonD850"])
    d = {}
    for name, v in zip(_FIT_COLS, row):
        if v is None:
            continue
        if name == "qbits":
            d["q"] = 1 / (2 ** v)
        else:
            d[name] = v
    return d


# (iso, Kmax, lam, sigGs, sigGssig, sigTL, sigTLsig, sigR, sigRsig, biassig) — process.py:260-289
_SONY_ISO = (
    (50, 0.047815, 0.1474653, 1.0164667, 0.005272454, 0.70727646, 0.004360543, 0.13997398, 0.0064381803, 0.010093017),
    (64, 0.0612032, 0.13243394, 1.0509665, 0.008081373, 0.71535635, 0.0056863446, 0.14346549, 0.006400559, 0.008690166),
    (80, 0.076504, 0.1121489, 1.180899, 0.011333668, 0.7799473, 0.009347968, 0.19540153, 0.008197397, 0.0107246125),
    (100, 0.09563, 0.14875287, 1.0067395, 0.0033682834, 0.70181876, 0.0037532174, 0.1391465, 0.006530218, 0.007235429),
    (125, 0.1195375, 0.12904578, 1.0279676, 0.007364685, 0.6961967, 0.0048687346, 0.14485553, 0.006731584, 0.008026363),
    (160, 0.153008, 0.094135, 1.1293099, 0.008340453, 0.7258587, 0.008032158, 0.19755602, 0.0082754735, 0.0101351),
    (200, 0.19126, 0.07902429, 1.2926387, 0.012171176, 0.8117464, 0.010250768, 0.22815849, 0.010726711, 0.011413908),
    (250, 0.239075, 0.051688068, 1.4345995, 0.01606571, 0.8630922, 0.013844714, 0.26271912, 0.0130637, 0.013569083),
    (320, 0.306016, 0.040700804, 1.7481371, 0.019626873, 1.0334468, 0.017629284, 0.3097104, 0.016202712, 0.017825918),
    (400, 0.38252, 0.0222538, 2.0595572, 0.024872316, 1.1816813, 0.02505812, 0.36209714, 0.01994737, 0.021005306),
    (500, 0.47815, -0.0031342343, 2.3956928, 0.030144656, 1.31772, 0.028629242, 0.42528257, 0.025104137, 0.02981831),
    (640, 0.612032, 0.002566592, 2.9662898, 0.045661453, 1.6474211, 0.04671843, 0.48839623, 0.031589635, 0.10000693),
    (800, 0.76504, -0.008199721, 3.5475867, 0.052318197, 1.9346539, 0.046128694, 0.5723769, 0.037824076, 0.025339302),
    (1000, 0.9563, -0.021061005, 4.2727833, 0.06972333, 2.2795107, 0.059203167, 0.6845563, 0.04879781, 0.027911892),
    (1250, 1.195375, -0.032423194, 5.177596, 0.092677385, 2.708437, 0.07622563, 0.8177013, 0.06162229, 0.03293372),
    (1600, 1.53008, -0.0441045, 6.29925, 0.1153261, 3.2283993, 0.09118158, 0.988786, 0.078567736, 0.03877672),
    (2000, 1.9126, -0.012963797, 2.653871, 0.015890995, 1.4356787, 0.02178686, 0.33124214, 0.018801652, 0.01570677),
    (2500, 2.39075, -0.027097283, 3.200225, 0.019307792, 1.6897862, 0.025873765, 0.38264316, 0.023769397, 0.018728448),
    (3200, 3.06016, -0.034863412, 3.9193838, 0.02649232, 2.0417721, 0.032873377, 0.44543457, 0.030114045, 0.021355819),
    (4000, 3.8252, -0.043700505, 4.8015847, 0.03781628, 2.4629273, 0.042401053, 0.52347374, 0.03929801, 0.026152484),
    (5000, 4.7815, -0.053150143, 5.8995814, 0.0625814, 2.9761007, 0.061326735, 0.6190265, 0.05335372, 0.058574405),
    (6400, 6.12032, -0.07517104, 7.1163535, 0.08435366, 3.4502964, 0.08226275, 0.7218788, 0.0642334, 0.059074216),
    (8000, 7.6504, -0.08208357, 8.916516, 0.12763213, 4.269624, 0.13381928, 0.87760293, 0.07389065, 0.084842026),
    (10000, 9.563, -0.073289566, 11.291476, 0.1639773, 5.495318, 0.16279395, 1.0522343, 0.094359785, 0.107438326),
    (12800, 12.24064, -0.06495205, 14.245901, 0.17283991, 7.038261, 0.18822834, 1.2749791, 0.120479785, 0.0944684),
    (16000, 15.3008, -0.060692135, 17.833515, 0.19809262, 8.877547, 0.23338738, 1.5559287, 0.15791349, 0.09725099),
    (20000, 19.126, -0.060213074, 22.084776, 0.21820943, 11.002351, 0.28806436, 1.8810822, 0.18937257, 0.4984733),
    (25600, 24.48128, -0.09089118, 25.853043, 0.35371417, 12.175712, 0.4215717, 2.2760193, 0.2609267, 0.37568903),
)


def get_specific_noise_params(camera_type=None, iso="100"):
    """process.py:257-308 — per-ISO calibrated point parameters; None for unknown cameras."""
    iso = str(iso)
    if camera_type == "SonyA7S2":
        for r in _SONY_ISO:
            if str(r[0]) == iso:
                return {"Kmax": r[1], "lam": r[2], "sigGs": r[3], "sigGssig": r[4], "sigTL": r[5],
                        "sigTLsig": r[6], "sigR": r[7], "sigRsig": r[8], "bias": 0, "biassig": r[9],
                        "q": 6.103515625e-05, "wp": 16383, "bl": 512}
        raise KeyError(iso)
    if camera_type == "IMX686":
        if iso == "100":
            return {"Kmax": 0.083805, "sigGs": 0.6926457, "sigGssig": 0.002096, "sigTL": 0.67998, "lam": 0.015,
                    "sigR": 0.23668, "q": 1 / (2 ** 10), "wp": 1023, "bl": 64, "bias": np.array([0, 0, 0, 0])}
        if iso == "6400":
            return {"Kmax": 8.74253, "sigGs": 14.30362, "sigGssig": 0.06967, "sigTL": 12.8901, "lam": 0.015,
                    "sigR": 0, "q": 1 / (2 ** 10), "wp": 1023, "bl": 64,
                    "bias": np.array([-0.08113494, -0.04906388, -0.9408157, -1.2048522])}
        raise KeyError(iso)
    return None


# --------------------------------------------------------------------------------------
# S2 / S3  parameter sampling (NumPy global RandomState, same draw order as the reference)
# --------------------------------------------------------------------------------------
_DUAL_ISO = ("SonyA7S2",)


def sample_params(camera_type="NikonD850", ln_ratio=False):
    """process.py:354-412.  Draw order: randint(2) [dual-ISO] → uniform(Kmin,Kmax) →
    normal ×4 (TL, R, Gs, bias) → uniform ratio.  IMX686 / NikonD850 raise KeyError('uReadk')
    exactly like the reference (:392 is unguarded)."""
    rs = np.random
    if camera_type in _DUAL_ISO:
        camera_type += "_lowISO" if rs.randint(2) < 1 else "_highISO"
    P = get_camera_noisy_params(camera_type)
    wp, bl, lam, q = P["wp"], P["bl"], P["lam"], P["q"]
    if camera_type in ("CRVD", "BM3D"):
        a_list = np.array([3.513262, 6.955588, 13.486051, 26.585953, 52.032536])
        b_list = np.array([11.917691, 38.117816, 130.818508, 484.539790, 1819.818657])
        bias_points = np.array([-1.12660, -1.69546, -3.25935, -6.68111, -12.66876])
        c = rs.randint(5)
        log_K = np.log(a_list)[c]
        K = a_list[c]
        mu_TL = P["sigTLk"] * log_K + P["sigTLb"] if "sigTLk" in P else 0
        mu_R = P["sigRk"] * log_K + P["sigRb"] if "sigRk" in P else 0
        mu_Gs = np.log(np.sqrt(b_list))[c]
        bias = bias_points[c]
    else:
        log_K = rs.uniform(low=P["Kmin"], high=P["Kmax"])
        K = np.exp(log_K)
        mu_TL = P["sigTLk"] * log_K + P["sigTLb"] if "sigTLk" in P else q
        mu_R = P["sigRk"] * log_K + P["sigRb"] if "sigRk" in P else q
        mu_Gs = P["sigGsk"] * log_K + P["sigGsb"] if "sigGsk" in P else q
        mu_bias = P["uReadk"] * log_K + P["uReadb"]          # KeyError for cameras without uRead*
    log_sigTL = rs.normal(loc=mu_TL, scale=P["sigTLsig"]) if "sigTLk" in P else 0
    log_sigR = rs.normal(loc=mu_R, scale=P["sigRsig"]) if "sigRk" in P else 0
    log_sigGs = rs.normal(loc=mu_Gs, scale=P["sigGssig"]) if "sigGsk" in P else q
    log_bias = rs.normal(loc=mu_bias, scale=P["uReadsig"]) if "uReadk" in P else 0
    sigTL, sigR, sigGs, bias = np.exp(log_sigTL), np.exp(log_sigR), np.exp(log_sigGs), np.exp(log_bias)
    if ln_ratio:
        high = 1 if "CRVD" in camera_type else 5
        ratio = np.exp(rs.uniform(low=-0.01, high=high))
    else:
        ratio = rs.uniform(low=100, high=300)
    return {"K": K, "sigTL": sigTL, "sigR": sigR, "sigGs": sigGs, "bias": bias,
            "lam": lam, "q": q, "ratio": ratio, "wp": wp, "bl": bl}


def sample_params_max(camera_type="NikonD850", ratio=None, iso=None):
    """process.py:311-351."""
    rs = np.random
    P = None
    if iso is not None:
        P = get_specific_noise_params(camera_type=camera_type, iso=iso)
    if P is None:
        if camera_type in _DUAL_ISO:
            camera_type += "_lowISO" if rs.randint(2) < 1 else "_highISO"
        P = get_camera_noisy_params(camera_type)
        bias = 0
        log_K = P["Kmax"] + rs.uniform(low=-0.01, high=+0.01)
        K = np.exp(log_K)
        mu_TL = P["sigTLk"] * log_K + P["sigTLb"]
        mu_R = P["sigRk"] * log_K + P["sigRb"]
        mu_Gs = P["sigGsk"] * log_K + P["sigGsb"] if "sigGsk" in P else 2 ** (-14)
        sigTL = np.exp(mu_TL)
        sigR = np.exp(mu_R)
        sigGs = np.exp(rs.normal(loc=mu_Gs, scale=P["sigGssig"]) if "sigGssig" in P else mu_Gs)
    else:
        K = P["Kmax"] * (1 + rs.uniform(low=-0.01, high=+0.01))
        sigGs = rs.normal(loc=P["sigGs"], scale=P["sigGssig"]) if "sigGssig" in P else P["sigGs"]
        sigTL = rs.normal(loc=P["sigTL"], scale=P["sigTLsig"]) if "sigTLsig" in P else P["sigTL"]
        sigR = rs.normal(loc=P["sigR"], scale=P["sigRsig"]) if "sigRsig" in P else P["sigR"]
        bias = P["bias"]
    wp, bl, lam, q = P["wp"], P["bl"], P["lam"], P["q"]
    if ratio is None:
        if "SonyA7S2" in camera_type:
            ratio = rs.uniform(low=100, high=300)
        else:
            ratio = np.exp(rs.uniform(low=0, high=2.08))
    return {"K": K, "sigTL": sigTL, "sigR": sigR, "sigGs": sigGs, "bias": bias,
            "lam": lam, "q": q, "ratio": ratio, "wp": wp, "bl": bl}


# --------------------------------------------------------------------------------------
# N1-N3  generate_noisy_obs, split into (a) the draws and (b) the deterministic arithmetic
# --------------------------------------------------------------------------------------

def parse_noise_code(noise_code: str) -> dict:
    """process.py:598-603."""
    c = noise_code.lower()
    return {k: (ch in c) for k, ch in (("P", "p"), ("TL", "g"), ("R", "r"), ("Q", "q"), ("D", "d"), ("black", "b"))}


def tukeylambda_ppf(u, lam):
    """SciPy's generic inverse-CDF sampling for tukeylambda (third-party, scipy 1.18.1,
    stats/_continuous_distns.py tukeylambda_gen._ppf): boxcox(u, lam) - boxcox1p(-u, lam), i.e.
    (u^lam - 1)/lam - ((1-u)^lam - 1)/lam.  We call the same scipy.special kernels so the float64
    value — and therefore its float32 cast — is the one the reference gets."""
    from scipy import special as sc
    u = np.asarray(u, dtype=F64)
    return sc.boxcox(u, lam) - sc.boxcox1p(-u, lam)


def is_f64_scalar(x) -> bool:
    """NEP-50: np.float64 scalars are 'strong' (promote a float32 array to float64);
    python floats / ints are weak."""
    return isinstance(x, (np.floating, np.ndarray)) and np.asarray(x).dtype == np.float64


def lam_of(y, p):
    """The Poisson rate exactly as the reference forms it (process.py:593-595,606):
    float32 scale-in, then division by K in K's precision."""
    y = np.asarray(y, dtype=F32)
    y = y * (p["wp"] - p["bl"])
    y = y / p["ratio"]
    return y, (1.0 * y / p["K"])


def draw_reference_order(shape, p, noise_code, rng=None):
    """Replays the RandomState draw order of generate_noisy_obs (process.py:605-616):
    poisson(shape) | randn(shape) → uniform(shape) [Tukey, inside scipy rvs] |
    standard_normal(shape) [Gaussian read] → randn(c,h,1) → uniform(-.5,.5,shape).
    Returns dict(counts|shot_z, read, row_z, q) with the dtypes the reference holds them in.
    `lam` must be supplied through p['_lam'] (float64 array)."""
    rs = np.random if rng is None else rng
    f = parse_noise_code(noise_code)
    d = {}
    lam = p["_lam"]
    if f["P"]:
        d["counts"] = rs.poisson(lam).astype(F32)
    else:
        d["shot_z"] = rs.randn(*shape).astype(F32)
    if not f["black"]:
        if f["TL"]:
            u = rs.uniform(size=shape)                       # scipy rvs → _ppf(U) * scale + loc
            d["read"] = (tukeylambda_ppf(u, p["lam"]) * (p["sigTL"] / 1.0) + 0).astype(F32)
        else:
            d["read"] = (rs.standard_normal(size=shape) * (p["sigGs"] / 1.0) + 0).astype(F32)
        if f["R"]:
            d["row_z"] = rs.randn(shape[-3], shape[-2], 1).astype(F32)
        if f["Q"]:
            d["q"] = rs.uniform(low=-0.5, high=0.5, size=shape)   # float64, in DN (param['q'] unused)
    return d


def noisy_obs_tail(y, p, noise_code, draws, ori=False, clip=False):
    """Deterministic arithmetic of generate_noisy_obs (process.py:593-631) given the draws,
    written with *explicit* dtypes instead of relying on NumPy promotion:

      chain "f64"  — K / sig* are np.float64 (what sample_params returns): every term after the
                     float32 casts is float64 and the result is rounded to float32 once.
      chain "weak" — K / sig* are python floats (sample_params_max(iso=...)): float32 throughout,
                     except that the float64 quantisation array promotes the running sum.
    """
    f = parse_noise_code(noise_code)
    y32, _ = lam_of(y, p)
    K, sigR, ratio = p["K"], p["sigR"], p["ratio"]
    span = p["wp"] - p["bl"]
    MFM = 1.0
    if f["P"]:
        shot = draws["counts"] * K / MFM
    else:
        shot = y32 + draws["shot_z"] * np.sqrt(np.maximum(y32 / K, 1e-10)) * K / MFM
    if not f["black"]:
        read = draws["read"]
        row = draws["row_z"] * sigR / MFM if f["R"] else 0
        q = draws["q"] if f["Q"] else 0
        bias = p["bias"].reshape(-1, 1, 1) if f["D"] else 0   # AttributeError for python-scalar bias, as the reference
    else:
        read = row = q = bias = 0
    z = (shot + read + row + q + bias) / span
    z = np.clip(z, -p["bl"] / p["wp"], 1) if not clip else np.clip(z, 0, 1)
    if ori is False:
        z = z * ratio
    return np.asarray(z).astype(F32)


def generate_noisy_obs(y, camera_type=None, wp=16383, noise_code="p", param=None, MultiFrameMean=1,
                       ori=False, clip=False, return_draws=False):
    """process.py:591-631 restated as draws + tail (MultiFrameMean is 1 at every call site)."""
    assert MultiFrameMean == 1
    p = dict(param)
    y = np.asarray(y, dtype=F32)
    _, lam = lam_of(y, p)
    p["_lam"] = lam
    draws = draw_reference_order(y.shape, p, noise_code)
    z = noisy_obs_tail(y, p, noise_code, draws, ori=ori, clip=clip)
    return (z, draws) if return_draws else z


# --------------------------------------------------------------------------------------
# N4  generate_noisy_torch (float32 chain)              data_process/process.py:634-673
# --------------------------------------------------------------------------------------

def noisy_torch_tail(y, p, noise_code, draws, ori=False, clip=False):
    """Arithmetic of generate_noisy_torch in float32, torch op order, given draws:
    counts (Poisson sample, f32), read (already-scaled N(0, sigGs) sample, f32),
    row_z (standard normal (c,h,1)), q_u (uniform [0,1) f32)."""
    f = parse_noise_code(noise_code)
    g = lambda v: F32(v)
    y = np.asarray(y, dtype=F32)
    wp, bl, K, ratio = g(p["wp"]), g(p["bl"]), g(p["K"]), g(p["ratio"])
    span = F32(wp - bl)
    y = y * span
    y = y / ratio
    MFM = F32(1.0)
    if not f["P"]:
        raise TypeError("reference generate_noisy_torch fails without 'p' (process.py:651)")
    shot = draws["counts"].astype(F32) * K / MFM
    if f["black"]:
        read = F32(0)
    else:
        if f["TL"]:
            raise NotImplementedError
        read = draws["read"].astype(F32)
    row = draws["row_z"].astype(F32) * g(p["sigR"]) / MFM if f["R"] else F32(0)
    q = (draws["q_u"].astype(F32) - F32(0.5)) * g(p["q"]) * span if f["Q"] else F32(0)
    if f["D"]:
        raise TypeError("reference generate_noisy_torch fails with 'd' (process.py:663)")
    z = (shot + read + row + q) / span
    lo = F32(-bl / wp)
    z = np.clip(z, lo, F32(1)) if not clip else np.clip(z, F32(0), F32(1))
    if ori is False:
        z = z * ratio
    return z.astype(F32)


# --------------------------------------------------------------------------------------
# D1 / D2  dataset-side glue                            data_process/syn_datasets.py
# --------------------------------------------------------------------------------------

def data_aug(data, mode=0):
    """syn_datasets.py:100-107 — rot90 by mode%4 on the last two axes, W-flip if mode//4."""
    if mode == 0:
        return data
    data = np.rot90(data, k=mode % 4, axes=(-2, -1))
    if mode // 4:
        data = data[..., ::-1]
    return data


def random_crop(img, h_start, w_start, patch, aug):
    """syn_datasets.py:162-173 with the crop points passed in: crop_per_image = len(aug) crops from the first crop points
    (IndexError when there are fewer points, as the reference's h_start[i])."""
    c = img.shape[0]
    crops = np.empty((len(aug), c, patch, patch), dtype=F32)
    for i in range(len(aug)):
        hs, ws = h_start[i], w_start[i]
        crops[i] = data_aug(img[:, hs:hs + patch, ws:ws + patch], mode=aug[i])
    return crops


def eval_crop(data, patch, base=64):
    """syn_datasets.py:109-133 — (1,c,h,w) -> (nh*nw, c, patch, patch): reflect-pad by base/2, tiles every l = patch - base,
    the last row / column of tiles flush with the padded frame's far edge."""
    _, c, h, w = data.shape
    d, l = base // 2, patch - base
    nh, nw = h // l + 1, w // l + 1
    pad = np.pad(data, ((0, 0), (0, 0), (d, d), (d, d)), mode="reflect")
    out = np.empty((nh, nw, c, patch, patch), dtype=data.dtype)
    for i in range(nh - 1):
        for j in range(nw - 1):
            out[i, j] = pad[0, :, i * l:i * l + patch, j * l:j * l + patch]
    for i in range(nh - 1):
        out[i, nw - 1] = pad[0, :, i * l:i * l + patch, -patch:]
    for j in range(nw - 1):
        out[nh - 1, j] = pad[0, :, -patch:, j * l:j * l + patch]
    out[nh - 1, nw - 1] = pad[0, :, -patch:, -patch:]
    return out.reshape(-1, c, patch, patch)


def eval_merge(tiles, h, w, base=64):
    """syn_datasets.py:135-159 — the inverse: interior l x l of every tile, written in the reference's order."""
    n, c, patch, _ = tiles.shape
    d, l = base // 2, patch - base
    nh, nw = h // l + 1, w // l + 1
    t = tiles.reshape(nh, nw, c, patch, patch)
    out = np.empty((1, c, h, w), dtype=tiles.dtype)
    for i in range(nh - 1):
        for j in range(nw - 1):
            out[..., i * l:i * l + l, j * l:j * l + l] = t[i, j, :, d:-d, d:-d]
    for i in range(nh - 1):
        out[..., i * l:i * l + l, -l:] = t[i, nw - 1, :, d:-d, d:-d]
    for j in range(nw - 1):
        out[..., -l:, j * l:j * l + l] = t[nh - 1, j, :, d:-d, d:-d]
    out[..., -l:, -l:] = t[nh - 1, nw - 1, :, d:-d, d:-d]
    return out


def darkshading_raw2bayer(lr_raw, darkshading, wp=16383, bl=512, add_mean=False, bias_draw=None, clip=False):
    """data_process/real_datasets.py:360-372 -> raw2bayer(norm=True, clip=False): NumPy's own promotion decides the
    precision (uint16 - float32 map -> float32; - float64 map -> float64; in-place += keeps the array dtype)."""
    lr = lr_raw - darkshading
    if add_mean:
        lr = lr + darkshading.mean()
    if bias_draw is not None:
        lr += bias_draw
    return raw2bayer(lr, wp=wp, bl=bl, norm=True, clip=clip)


def hbr_map(data, lut, rand, norm=True, keep_remainder=True):
    """HighBitRecovery.map (data_process/process.py:726-751) with the uniforms passed in; `lut` is the dict HB2LB_LUT
    returns (keys param, dist, low, high, and per integer level cdf / range).  keep_remainder = the object's `float` flag."""
    p = lut['param']
    if np.max(data) <= 1:
        data = data * (p['wp'] - p['bl'])
    data_float = data.copy()
    data = np.round(data_float)
    delta = data_float - data
    for x in range(lut['low'], lut['high']):
        keys = (data == x)
        data[keys] = lut['dist'].ppf(lut[x]['cdf'] + rand[keys] * lut[x]['range'])
    if keep_remainder:
        data = data + delta
    return data / (p['wp'] - p['bl']) if norm else data + p['bl']


def random_gains(camera_type="SonyA7S2"):
    """data_process/unprocess.py:60-77: (rgb_gain, red_gain, blue_gain) as float32 arrays of shape (1,).
    Draw order: one torch.distributions.Normal(0.8, 0.1) sample (torch's global CPU generator), then one
    np.random.uniform for the red gain; the blue gain is a quadratic fit of the red gain (float64, rounded to float32 once)."""
    import torch
    import torch.distributions as tdist
    n = tdist.Normal(loc=torch.tensor([0.8]), scale=torch.tensor([0.1]))
    rgb_gain = 1.0 / n.sample()
    if camera_type == "SonyA7S2":
        red_gain = np.random.uniform(1.75, 2.65)
        fit = [14.65, -9.63942308, 1.80288462]
    elif camera_type == "IMX686":
        red_gain = np.random.uniform(1.4, 2.3)
        fit = [6.14381188, -3.65620261, 0.70205967]
    else:
        raise NotImplementedError
    blue_gain = fit[0] + fit[1] * red_gain + fit[2] * red_gain ** 2
    return rgb_gain.numpy(), np.array([red_gain]).astype(F32), np.array([blue_gain]).astype(F32)


def wb_jitter(hr_crops, wb, gains):
    """syn_datasets.py:313-319 given the gains of random_gains(): every plane times rgb_gain (float32, in place), then plane 0
    times wb[0] / red_gain and plane 2 times wb[2] / blue_gain.  The type of `wb` decides the arithmetic of the second product
    (NEP 50): a float32 / python-float white balance keeps it float32; an np.float64 white balance makes `wb / gain` a float64
    array, the product float64, and the assignment rounds it to float32 once."""
    rgb_gain, red_gain, blue_gain = gains
    hr_crops = hr_crops.copy()
    red = wb[0] / red_gain
    blue = wb[2] / blue_gain
    hr_crops *= rgb_gain
    hr_crops[:, 0] = hr_crops[:, 0] * red
    hr_crops[:, 2] = hr_crops[:, 2] * blue
    return hr_crops


def post_synth_clip(lr, hr, clip):
    """syn_datasets.py:339-342 / trainer_SID.py:481-485.  clip==2 (HALF_CLIP) → lower bound -inf."""
    if clip:
        lb = -np.inf if clip == 2 else 0
        lr = lr.clip(lb, 1)
        hr = hr.clip(0, 1)
    return lr, hr


# --------------------------------------------------------------------------------------
# U1-U3  networks (torch CPU fp32 functional restatement)
# --------------------------------------------------------------------------------------

def unet_forward(x, sd, res=False):
    """archs/Unet.py:54-99 as torch.nn.functional calls on a reference-keyed state_dict."""
    import torch
    import torch.nn.functional as Fn
    act = lambda t: Fn.leaky_relu(t, 0.2)
    cv = lambda t, n: Fn.conv2d(t, sd[n + ".weight"], sd[n + ".bias"], padding=sd[n + ".weight"].shape[-1] // 2)
    up = lambda t, n: Fn.conv_transpose2d(t, sd[n + ".weight"], sd[n + ".bias"], stride=2)
    c1 = act(cv(act(cv(x, "conv1_1")), "conv1_2"))
    c2 = act(cv(act(cv(Fn.max_pool2d(c1, 2), "conv2_1")), "conv2_2"))
    c3 = act(cv(act(cv(Fn.max_pool2d(c2, 2), "conv3_1")), "conv3_2"))
    c4 = act(cv(act(cv(Fn.max_pool2d(c3, 2), "conv4_1")), "conv4_2"))
    c5 = act(cv(act(cv(Fn.max_pool2d(c4, 2), "conv5_1")), "conv5_2"))
    c6 = act(cv(act(cv(torch.cat([up(c5, "upv6"), c4], 1), "conv6_1")), "conv6_2"))
    c7 = act(cv(act(cv(torch.cat([up(c6, "upv7"), c3], 1), "conv7_1")), "conv7_2"))
    c8 = act(cv(act(cv(torch.cat([up(c7, "upv8"), c2], 1), "conv8_1")), "conv8_2"))
    c9 = act(cv(act(cv(torch.cat([up(c8, "upv9"), c1], 1), "conv9_1")), "conv9_2"))
    out = cv(c9, "conv10_1")
    return out + x if res else out


def resunet_forward(x, sd, res=False):
    """archs/ResUnet.py:46-88 + archs/modules.py:130-197.  ResidualBlock(is_activate=False):
    conv(no bias)+ReLU → conv(no bias); += shortcut (identity, or bias-free 1×1 when in≠out).
    Down-sampling is a stride-2 3×3 conv WITH bias and NO activation (the add_module('relu')
    at modules.py:134-135 hangs on an nn.Conv2d and never runs)."""
    import torch
    import torch.nn.functional as Fn

    def block(t, n):
        o = Fn.relu(Fn.conv2d(t, sd[f"{n}.block.0.conv.conv.weight"], None, padding=1))
        o = Fn.conv2d(o, sd[f"{n}.block.1.conv.conv.weight"], None, padding=1)
        k = f"{n}.short_cut.0.conv.conv.weight"
        return o + (Fn.conv2d(t, sd[k], None) if k in sd else t)

    down = lambda t, n: Fn.conv2d(t, sd[f"{n}.conv.weight"], sd[f"{n}.conv.bias"], stride=2, padding=1)
    up = lambda t, n: Fn.conv_transpose2d(t, sd[n + ".weight"], sd[n + ".bias"], stride=2)
    cin = Fn.relu(Fn.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1))
    c1 = block(cin, "conv1")
    c2 = block(down(c1, "pool1"), "conv2")
    c3 = block(down(c2, "pool2"), "conv3")
    c4 = block(down(c3, "pool3"), "conv4")
    c5 = block(down(c4, "pool4"), "conv5")
    c6 = block(torch.cat([up(c5, "upv6"), c4], 1), "conv6")
    c7 = block(torch.cat([up(c6, "upv7"), c3], 1), "conv7")
    c8 = block(torch.cat([up(c7, "upv8"), c2], 1), "conv8")
    c9 = block(torch.cat([up(c8, "upv9"), c1], 1), "conv9")
    out = Fn.conv2d(c9, sd["conv10.weight"], sd["conv10.bias"])
    return out + x if res else out


# --------------------------------------------------------------------------------------
# E1 / E2  eval boundary and metrics
# --------------------------------------------------------------------------------------

def illuminance_correct(predict, source):
    """data_process/__init__.py:162-175 (N==1): clamp → <p,s>/<p,p> over source!=1 → scale."""
    import torch
    predict = torch.clamp(predict, 0, 1)
    m = source != 1
    pc, sc = predict[m], source[m]
    return torch.dot(pc, sc) / torch.dot(pc, pc) * predict


def tensor2im(t):
    """utils/visualization.py:9-24 — NCHW[0] → HWC, ×255, clip; no rounding."""
    a = np.asarray(t[0], dtype=F32)
    return np.clip(np.transpose(a, (1, 2, 0)) * 255.0, 0, 255)


def psnr(target, estimate, data_range=255):
    """skimage.metrics.peak_signal_noise_ratio (documented defaults): float64 MSE."""
    err = np.mean((np.asarray(target, F64) - np.asarray(estimate, F64)) ** 2, dtype=F64)
    return 10 * np.log10((data_range ** 2) / err)


def ssim(target, estimate, data_range=255):
    """skimage.metrics.structural_similarity(channel_axis=-1) documented defaults: 7×7 uniform
    window, K1=0.01, K2=0.03, sample covariance (N/(N-1)), border of 3 cropped, mean over
    pixels then over channels; float64 (inputs are float32 → skimage promotes per _supported_float_type
    to float32 for the filters; we use float64 and compare with a tolerance)."""
    from scipy.ndimage import uniform_filter
    X = np.asarray(target, F64)
    Y = np.asarray(estimate, F64)
    win, K1, K2 = 7, 0.01, 0.03
    NP = win * win
    cov_norm = NP / (NP - 1)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    vals = []
    for ch in range(X.shape[-1]):
        x, y = X[..., ch], Y[..., ch]
        ux, uy = uniform_filter(x, win), uniform_filter(y, win)
        uxx, uyy, uxy = uniform_filter(x * x, win), uniform_filter(y * y, win), uniform_filter(x * y, win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        pad = (win - 1) // 2
        vals.append(S[pad:-pad, pad:-pad].mean(dtype=F64))
    return float(np.mean(vals))


def ssim_float32_path(target, estimate, data_range=255):
    """The same algorithm the way scikit-image >= 0.19 executes it for float32 images (`_supported_float_type` keeps float32: the
    five uniform filters, their products and the S map are float32; only the final mean is float64).  scikit-image is not in this
    image's wheelhouse, so neither restatement can be checked against the library itself (E2 stays "parity unpinned"); this one
    bounds what the library's own float32 rounding could move: tests require |ssim - ssim_float32_path| < 5e-5."""
    from scipy.ndimage import uniform_filter
    X = np.asarray(target, F32)
    Y = np.asarray(estimate, F32)
    win, K1, K2 = 7, 0.01, 0.03
    cov_norm = (win * win) / (win * win - 1)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    vals = []
    for ch in range(X.shape[-1]):
        x, y = X[..., ch], Y[..., ch]
        ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
        uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        A1, A2, B1, B2 = 2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2
        S = (A1 * A2) / (B1 * B2)
        assert S.dtype == F32
        pad = (win - 1) // 2
        vals.append(S[pad:-pad, pad:-pad].mean(dtype=F64))
    return float(np.mean(vals))


def l1_loss(pred, target):
    """losses/base_loss.py:92-103 via trainer_SID.py:99 — F.l1_loss(pred.clamp(0,1), hr)."""
    return float(np.mean(np.abs(np.clip(pred, 0, 1) - target)))


# --------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11) — the counter-based generator the CUDA kernels use.
# Restated here so tests can predict the device draws bit-for-bit on the CPU.
# --------------------------------------------------------------------------------------
_PH_M0, _PH_M1 = 0xD2511F53, 0xCD9E8D57
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """ctr: (...,4) uint32, key: (2,) uint32 → (...,4) uint32."""
    c = np.asarray(ctr, dtype=np.uint64).copy()
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0 = _PH_M0 * c[..., 0]
        p1 = _PH_M1 * c[..., 2]
        hi0, lo0 = p0 >> 32, p0 & 0xFFFFFFFF
        hi1, lo1 = p1 >> 32, p1 & 0xFFFFFFFF
        n0 = hi1 ^ c[..., 1] ^ k0
        n2 = hi0 ^ c[..., 3] ^ k1
        c = np.stack([n0, lo1, n2, lo0], axis=-1) & 0xFFFFFFFF
        k0 = (k0 + _PH_W0) & 0xFFFFFFFF
        k1 = (k1 + _PH_W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


# --------------------------------------------------------------------------------------
# The same arithmetic with every cast spelled out — this is the specification the CUDA
# `noise_tail` device function is written from (csrc/noise_core.cuh).  Tested equal to
# noisy_obs_tail (and hence to the live reference) in tests/test_oracle.py.
# --------------------------------------------------------------------------------------

def chain_flags(p) -> dict:
    """Which scalars are 'strong' float64 under NEP-50 for this param dict."""
    return {"k64": is_f64_scalar(p["K"]), "ratio64": is_f64_scalar(p["ratio"]),
            "sig64": is_f64_scalar(p["sigR"])}


def noisy_obs_tail_explicit(y, p, noise_code, draws, ori=False, clip=False):
    f = parse_noise_code(noise_code)
    cf = chain_flags(p)
    k64, ratio64, sig64 = cf["k64"], cf["ratio64"], cf["sig64"]
    span = p["wp"] - p["bl"]
    K64, K32 = F64(p["K"]), F32(p["K"])
    y32 = (np.asarray(y, F32) * F32(span)).astype(F32)
    if ratio64:
        ysc, y_is64 = y32.astype(F64) / F64(p["ratio"]), True
    else:
        ysc, y_is64 = (y32 / F32(p["ratio"])).astype(F32), False

    def promote(a, a64, want64):
        return (a.astype(F64), True) if (want64 and not a64) else (a, a64)

    # ---- shot
    if f["P"]:
        cnt = draws["counts"].astype(F32)
        if k64:
            acc, a64 = cnt.astype(F64) * K64, True
        else:
            acc, a64 = (cnt * K32).astype(F32), False
    else:
        z0 = draws["shot_z"].astype(F32)
        if k64 or y_is64:
            kk = K64        # a weak python-float K meets a float64 array here: full double value
            t = np.sqrt(np.maximum(ysc.astype(F64) / kk, F64(1e-10)))
            acc, a64 = ysc.astype(F64) + z0.astype(F64) * t * kk, True
        else:
            t = np.sqrt(np.maximum((ysc / K32).astype(F32), F32(1e-10))).astype(F32)
            acc = (ysc + ((z0 * t).astype(F32) * K32).astype(F32)).astype(F32)
            a64 = False
    # ---- read / row / q / bias
    if not f["black"]:
        rd = draws["read"].astype(F32)
        acc = acc + (rd.astype(F64) if a64 else rd)
        if f["R"]:
            rz = draws["row_z"].astype(F32)
            if sig64:
                acc, a64 = promote(acc, a64, True)
                acc = acc + rz.astype(F64) * F64(p["sigR"])
            else:
                row = (rz * F32(p["sigR"])).astype(F32)
                acc = acc + (row.astype(F64) if a64 else row)
        if f["Q"]:
            acc, a64 = promote(acc, a64, True)
            acc = acc + draws["q"].astype(F64)
        if f["D"]:
            # bias is an ndarray or np.float64 here (python scalars raise in the reference):
            # int64 / float64 arrays both promote the float32 sum to float64
            b = p["bias"].reshape(-1, 1, 1)
            acc, a64 = promote(acc, a64, True)
            acc = acc + b.astype(F64)
    # ---- scale-out
    lo = -p["bl"] / p["wp"]
    if a64:
        z = acc / F64(span)
        z = np.clip(z, F64(0), F64(1)) if clip else np.clip(z, F64(lo), F64(1))
    else:
        z = (acc / F32(span)).astype(F32)
        z = np.clip(z, F32(0), F32(1)) if clip else np.clip(z, F32(lo), F32(1))
    if ori is False:
        if a64 or ratio64:
            z = z.astype(F64) * F64(p["ratio"])
        else:
            z = (z * F32(p["ratio"])).astype(F32)
    return z.astype(F32)
