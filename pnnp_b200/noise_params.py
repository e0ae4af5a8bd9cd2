"""Camera noise tables and parameter sampling (host scalars) behind the reference's names:
get_camera_noisy_params / get_specific_noise_params / sample_params / sample_params_max
(data_process/process.py:215-412).  Draw order on NumPy's global RandomState is the reference's,
so the same `np.random.seed` yields the same dict (tests/test_params.py checks the goldens).

Also: ParamTable — the packed device-side table (one 128-byte row per crop) the kernels read.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

Dual_ISO_Cameras = ["SonyA7S2"]
HALF_CLIP = 2

# log-linear fits ln(sigma) = k ln(K) + b with residual sigma (process.py:217-250)
_CAMERA_FITS = {
    "NikonD850": dict(Kmin=1.2, Kmax=2.4828, lam=-0.26, q=1 / (2 ** 14), wp=16383, bl=512,
                      sigTLk=0.906, sigTLb=-0.6754, sigTLsig=0.035165,
                      sigRk=0.8322, sigRb=-2.3326, sigRsig=0.301333,
                      sigGsk=0.8322, sigGsb=-0.1754, sigGssig=0.035165),
    "IMX686": dict(Kmin=-0.19118, Kmax=2.16820, lam=0.102, q=1 / (2 ** 10), wp=1023, bl=64,
                   sigTLk=0.85187, sigTLb=0.07991, sigTLsig=0.02921,
                   sigRk=0.87611, sigRb=-2.11455, sigRsig=0.03274,
                   sigGsk=0.85187, sigGsb=0.67991, sigGssig=0.02921),
    "SonyA7S2_lowISO": dict(Kmin=-1.67214, Kmax=0.42228, lam=-0.026, q=1 / (2 ** 14), wp=16383, bl=512,
                            sigRk=0.78782, sigRb=-0.34227, sigRsig=0.02832,
                            sigTLk=0.74043, sigTLb=0.86182, sigTLsig=0.00712,
                            sigGsk=0.82966, sigGsb=1.49343, sigGssig=0.00359,
                            sigReadk=0.82879, sigReadb=1.50601, sigReadsig=0.00362,
                            uReadk=0.01472, uReadb=0.01129, uReadsig=0.00034),
    "SonyA7S2_highISO": dict(Kmin=0.64567, Kmax=2.51606, lam=-0.025, q=1 / (2 ** 14), wp=16383, bl=512,
                             sigRk=0.62945, sigRb=-1.51040, sigRsig=0.02609,
                             sigTLk=0.74901, sigTLb=-0.12348, sigTLsig=0.00638,
                             sigGsk=0.82878, sigGsb=0.44162, sigGssig=0.00153,
                             sigReadk=0.82645, sigReadb=0.45061, sigReadsig=0.00156,
                             uReadk=0.00385, uReadb=0.00674, uReadsig=0.00039),
    "CRVD": dict(Kmin=1.31339, Kmax=3.95448, lam=0.015, q=1 / (2 ** 12), wp=4095, bl=240,
                 sigRk=0.93368, sigRb=-2.19692, sigRsig=0.02473,
                 sigGsk=0.95387, sigGsb=0.01552, sigGssig=0.00855,
                 sigTLk=0.95495, sigTLb=0.01618, sigTLsig=0.00790),
}

# SonyA7S2 per-ISO calibration (process.py:260-289); columns:
_SONY_COLS = ("Kmax", "lam", "sigGs", "sigGssig", "sigTL", "sigTLsig", "sigR", "sigRsig", "biassig")
_SONY_POINTS = {
    50: (0.047815, 0.1474653, 1.0164667, 0.005272454, 0.70727646, 0.004360543, 0.13997398, 0.0064381803, 0.010093017),
    64: (0.0612032, 0.13243394, 1.0509665, 0.008081373, 0.71535635, 0.0056863446, 0.14346549, 0.006400559, 0.008690166),
    80: (0.076504, 0.1121489, 1.180899, 0.011333668, 0.7799473, 0.009347968, 0.19540153, 0.008197397, 0.0107246125),
    100: (0.09563, 0.14875287, 1.0067395, 0.0033682834, 0.70181876, 0.0037532174, 0.1391465, 0.006530218, 0.007235429),
    125: (0.1195375, 0.12904578, 1.0279676, 0.007364685, 0.6961967, 0.0048687346, 0.14485553, 0.006731584, 0.008026363),
    160: (0.153008, 0.094135, 1.1293099, 0.008340453, 0.7258587, 0.008032158, 0.19755602, 0.0082754735, 0.0101351),
    200: (0.19126, 0.07902429, 1.2926387, 0.012171176, 0.8117464, 0.010250768, 0.22815849, 0.010726711, 0.011413908),
    250: (0.239075, 0.051688068, 1.4345995, 0.01606571, 0.8630922, 0.013844714, 0.26271912, 0.0130637, 0.013569083),
    320: (0.306016, 0.040700804, 1.7481371, 0.019626873, 1.0334468, 0.017629284, 0.3097104, 0.016202712, 0.017825918),
    400: (0.38252, 0.0222538, 2.0595572, 0.024872316, 1.1816813, 0.02505812, 0.36209714, 0.01994737, 0.021005306),
    500: (0.47815, -0.0031342343, 2.3956928, 0.030144656, 1.31772, 0.028629242, 0.42528257, 0.025104137, 0.02981831),
    640: (0.612032, 0.002566592, 2.9662898, 0.045661453, 1.6474211, 0.04671843, 0.48839623, 0.031589635, 0.10000693),
    800: (0.76504, -0.008199721, 3.5475867, 0.052318197, 1.9346539, 0.046128694, 0.5723769, 0.037824076, 0.025339302),
    1000: (0.9563, -0.021061005, 4.2727833, 0.06972333, 2.2795107, 0.059203167, 0.6845563, 0.04879781, 0.027911892),
    1250: (1.195375, -0.032423194, 5.177596, 0.092677385, 2.708437, 0.07622563, 0.8177013, 0.06162229, 0.03293372),
    1600: (1.53008, -0.0441045, 6.29925, 0.1153261, 3.2283993, 0.09118158, 0.988786, 0.078567736, 0.03877672),
    2000: (1.9126, -0.012963797, 2.653871, 0.015890995, 1.4356787, 0.02178686, 0.33124214, 0.018801652, 0.01570677),
    2500: (2.39075, -0.027097283, 3.200225, 0.019307792, 1.6897862, 0.025873765, 0.38264316, 0.023769397, 0.018728448),
    3200: (3.06016, -0.034863412, 3.9193838, 0.02649232, 2.0417721, 0.032873377, 0.44543457, 0.030114045, 0.021355819),
    4000: (3.8252, -0.043700505, 4.8015847, 0.03781628, 2.4629273, 0.042401053, 0.52347374, 0.03929801, 0.026152484),
    5000: (4.7815, -0.053150143, 5.8995814, 0.0625814, 2.9761007, 0.061326735, 0.6190265, 0.05335372, 0.058574405),
    6400: (6.12032, -0.07517104, 7.1163535, 0.08435366, 3.4502964, 0.08226275, 0.7218788, 0.0642334, 0.059074216),
    8000: (7.6504, -0.08208357, 8.916516, 0.12763213, 4.269624, 0.13381928, 0.87760293, 0.07389065, 0.084842026),
    10000: (9.563, -0.073289566, 11.291476, 0.1639773, 5.495318, 0.16279395, 1.0522343, 0.094359785, 0.107438326),
    12800: (12.24064, -0.06495205, 14.245901, 0.17283991, 7.038261, 0.18822834, 1.2749791, 0.120479785, 0.0944684),
    16000: (15.3008, -0.060692135, 17.833515, 0.19809262, 8.877547, 0.23338738, 1.5559287, 0.15791349, 0.09725099),
    20000: (19.126, -0.060213074, 22.084776, 0.21820943, 11.002351, 0.28806436, 1.8810822, 0.18937257, 0.4984733),
    25600: (24.48128, -0.09089118, 25.853043, 0.35371417, 12.175712, 0.4215717, 2.2760193, 0.2609267, 0.37568903),
}
_IMX686_POINTS = {
    100: dict(Kmax=0.083805, sigGs=0.6926457, sigGssig=0.002096, sigTL=0.67998, lam=0.015, sigR=0.23668,
              q=1 / (2 ** 10), wp=1023, bl=64, bias=(0, 0, 0, 0)),
    6400: dict(Kmax=8.74253, sigGs=14.30362, sigGssig=0.06967, sigTL=12.8901, lam=0.015, sigR=0,
               q=1 / (2 ** 10), wp=1023, bl=64, bias=(-0.08113494, -0.04906388, -0.9408157, -1.2048522)),
}


def get_camera_noisy_params(camera_type=None):
    """process.py:215-255.  Unknown cameras fall back to NikonD850 like the reference."""
    return dict(_CAMERA_FITS.get(camera_type, _CAMERA_FITS["NikonD850"]))


def get_specific_noise_params(camera_type=None, iso="100"):
    """process.py:257-308.  KeyError for an ISO that is not tabulated, None for other cameras."""
    if camera_type == "SonyA7S2":
        row = _SONY_POINTS[int(iso)]
        d = dict(zip(_SONY_COLS, row))
        d.update(bias=0, q=6.103515625e-05, wp=16383, bl=512)
        return d
    if camera_type == "IMX686":
        d = dict(_IMX686_POINTS[int(iso)])
        d["bias"] = np.array(d["bias"])
        return d
    return None


def sample_params_max(camera_type="NikonD850", ratio=None, iso=None):
    """process.py:311-351 — parameters at the camera's maximum gain (or a tabulated ISO)."""
    rs = np.random
    point = get_specific_noise_params(camera_type=camera_type, iso=iso) if iso is not None else None
    if point is None:
        if camera_type in Dual_ISO_Cameras:
            camera_type += "_lowISO" if rs.randint(2) < 1 else "_highISO"
        fit = get_camera_noisy_params(camera_type)
        bias = 0
        log_K = fit["Kmax"] + rs.uniform(low=-0.01, high=+0.01)
        K = np.exp(log_K)
        sigTL = np.exp(fit["sigTLk"] * log_K + fit["sigTLb"])
        sigR = np.exp(fit["sigRk"] * log_K + fit["sigRb"])
        mu_Gs = fit["sigGsk"] * log_K + fit["sigGsb"] if "sigGsk" in fit else 2 ** (-14)
        sigGs = np.exp(rs.normal(loc=mu_Gs, scale=fit["sigGssig"]) if "sigGssig" in fit else mu_Gs)
        src = fit
    else:
        K = point["Kmax"] * (1 + rs.uniform(low=-0.01, high=+0.01))
        sigGs = rs.normal(loc=point["sigGs"], scale=point["sigGssig"]) if "sigGssig" in point else point["sigGs"]
        sigTL = rs.normal(loc=point["sigTL"], scale=point["sigTLsig"]) if "sigTLsig" in point else point["sigTL"]
        sigR = rs.normal(loc=point["sigR"], scale=point["sigRsig"]) if "sigRsig" in point else point["sigR"]
        bias = point["bias"]
        src = point
    if ratio is None:
        if "SonyA7S2" in camera_type:
            ratio = rs.uniform(low=100, high=300)
        else:
            ratio = np.exp(rs.uniform(low=0, high=2.08))
    return {"K": K, "sigTL": sigTL, "sigR": sigR, "sigGs": sigGs, "bias": bias,
            "lam": src["lam"], "q": src["q"], "ratio": ratio, "wp": src["wp"], "bl": src["bl"]}


_CRVD_A = np.array([3.513262, 6.955588, 13.486051, 26.585953, 52.032536])
_CRVD_B = np.array([11.917691, 38.117816, 130.818508, 484.539790, 1819.818657])
_CRVD_BIAS = np.array([-1.12660, -1.69546, -3.25935, -6.68111, -12.66876])


def sample_params(camera_type="NikonD850", ln_ratio=False):
    """process.py:354-412.  Like the reference this raises KeyError('uReadk') for cameras whose
    fit has no uRead* entries (IMX686, NikonD850): line 392 there is unguarded."""
    rs = np.random
    if camera_type in Dual_ISO_Cameras:
        camera_type += "_lowISO" if rs.randint(2) < 1 else "_highISO"
    fit = get_camera_noisy_params(camera_type)
    q = fit["q"]
    has = lambda stem: (stem + "k") in fit
    if camera_type in ("CRVD", "BM3D"):
        pick = rs.randint(5)
        log_K, K = np.log(_CRVD_A)[pick], _CRVD_A[pick]
        mu = {"sigTL": fit["sigTLk"] * log_K + fit["sigTLb"] if has("sigTL") else 0,
              "sigR": fit["sigRk"] * log_K + fit["sigRb"] if has("sigR") else 0,
              "sigGs": np.log(np.sqrt(_CRVD_B))[pick]}
    else:
        log_K = rs.uniform(low=fit["Kmin"], high=fit["Kmax"])
        K = np.exp(log_K)
        mu = {s: (fit[s + "k"] * log_K + fit[s + "b"] if has(s) else q) for s in ("sigTL", "sigR", "sigGs")}
        mu["uRead"] = fit["uReadk"] * log_K + fit["uReadb"]
    draw = lambda s, default: rs.normal(loc=mu[s], scale=fit[s + "sig"]) if has(s) else default
    log_sigTL, log_sigR, log_sigGs, log_bias = draw("sigTL", 0), draw("sigR", 0), draw("sigGs", q), draw("uRead", 0)
    if ln_ratio:
        ratio = np.exp(rs.uniform(low=-0.01, high=1 if "CRVD" in camera_type else 5))
    else:
        ratio = rs.uniform(low=100, high=300)
    return {"K": K, "sigTL": np.exp(log_sigTL), "sigR": np.exp(log_sigR), "sigGs": np.exp(log_sigGs),
            "bias": np.exp(log_bias), "lam": fit["lam"], "q": q, "ratio": ratio, "wp": fit["wp"], "bl": fit["bl"]}


# --------------------------------------------------------------------------------------------
# Device parameter table
# --------------------------------------------------------------------------------------------

def _scalar(v) -> float:
    if isinstance(v, torch.Tensor):
        return float(v.item())
    return float(np.asarray(v).reshape(-1)[0]) if isinstance(v, np.ndarray) else float(v)


def _is_f64(v) -> bool:
    """NEP-50: np.float64 scalars / arrays are strong, python numbers are weak."""
    return isinstance(v, (np.floating, np.ndarray)) and np.asarray(v).dtype == np.float64


def fill_row(row: "_lib.NoiseParamsRow", p: dict, torch_chain: bool = False) -> None:
    wp, bl = _scalar(p["wp"]), _scalar(p["bl"])
    row.K, row.sigTL, row.sigGs, row.sigR = _scalar(p["K"]), _scalar(p["sigTL"]), _scalar(p["sigGs"]), _scalar(p["sigR"])
    row.lam, row.q, row.ratio = _scalar(p["lam"]), _scalar(p["q"]), _scalar(p["ratio"])
    row.span = wp - bl
    if torch_chain:
        # torch computes -bl/wp on float32 tensors (process.py:668)
        row.clip_lo = float(np.float32(-np.float32(bl)) / np.float32(wp))
        row.flags = 0
    else:
        row.clip_lo = -bl / wp
        row.flags = ((_lib.F_K64 if _is_f64(p["K"]) else 0) | (_lib.F_RATIO64 if _is_f64(p["ratio"]) else 0)
                     | (_lib.F_SIG64 if _is_f64(p["sigR"]) else 0))
    b = p.get("bias", 0)
    if isinstance(b, torch.Tensor):
        b = b.detach().cpu().numpy()
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    for i in range(4):
        row.bias[i] = float(b[i % b.size])


class ParamTable:
    """n × 128-byte rows in pinned host memory + a device copy (uploaded asynchronously)."""

    def __init__(self, params, device, torch_chain=False):
        params = list(params)
        self.n = len(params)
        rows = (_lib.NoiseParamsRow * self.n)()
        for r, p in zip(rows, params):
            fill_row(r, p, torch_chain)
        host = torch.frombuffer(bytearray(bytes(rows)), dtype=torch.uint8)
        if torch.cuda.is_available():
            host = host.pin_memory()
        self.host = host
        self.device = host.to(device, non_blocking=True)
        self.ratios = [float(r.ratio) for r in rows]
        # hint for the launcher (PNNP_CODE_UNIFORM_F64): sample_params-style float64 parameters in every row AND a Tukey-lambda
        # shape away from 0 (the specialised kernel has only the power form of the quantile; |lam| < 1e-3 needs the series)
        self.uniform_f64 = all(r.flags == (_lib.F_K64 | _lib.F_SIG64) and abs(r.lam) >= 1e-3 for r in rows)

    def data_ptr(self):
        return self.device.data_ptr()
