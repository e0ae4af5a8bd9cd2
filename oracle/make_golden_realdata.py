"""Golden vectors for the real-data side rows (dark shading + HighBitRecovery), from the UNMODIFIED reference
(build container only):   python oracle/make_golden_realdata.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    R = rh.load()
    P, ISP = R.process, R.isp_ops
    out = {}
    rs = np.random.RandomState(3)
    # ---- dark shading: the reference's statements (real_datasets.py:360-372) executed with its own raw2bayer
    H, W = 24, 32
    raw = rs.randint(400, 2000, size=(H, W)).astype(np.uint16)
    ds_k = (rs.rand(H, W) * 1e-4).astype(np.float32)
    ds_b = (rs.randn(H, W) * 0.7).astype(np.float32)
    iso = 3200
    for tag, ble in (("f32", 1.25), ("f64", np.float64(1.2512345678))):       # python float keeps float32; np.float64 promotes
        ds = ds_k * iso + ds_b + ble
        for mean in (False, True):
            for bd in (None, 0.3712345):
                lr = raw - ds
                if mean:
                    lr = lr + ds.mean()
                if bd is not None:
                    lr += bd
                out[f"ds_{tag}_m{int(mean)}_b{int(bd is not None)}"] = ISP.raw2bayer(lr, wp=16383, bl=512, norm=True, clip=False)
        out[f"ds_{tag}_map"] = ds
    out["ds_raw"] = raw
    # ---- HighBitRecovery: LUT + map with NumPy's global RandomState (process.py:675-751)
    for k, (cam, code, iso, shape) in enumerate((("SonyA7S2", "pgrq", 3200, (2, 4, 16, 24)), ("IMX686", "prq", 6400, (1, 4, 12, 20)))):
        np.random.seed(100 + k)
        hb = P.HighBitRecovery(camera_type=cam, noise_code=code)
        hb.get_lut([iso], blc_mean=None)
        lut = hb.lut[iso]
        p = lut["param"]
        span = p["wp"] - p["bl"]
        data = (rs.randn(*shape) * lut["sigma"] * 1.5 / span).astype(np.float32)       # dark frame, normalised
        np.random.seed(200 + k)
        res = hb.map(data.copy(), iso, norm=True)
        np.random.seed(200 + k)
        rand = np.random.uniform(0, 1, size=data.shape)
        out[f"hbr{k}_data"], out[f"hbr{k}_rand"], out[f"hbr{k}_out"] = data, rand, res
        out[f"hbr{k}_lut"] = np.array([lut["low"], lut["high"], lut["bias"], lut["sigma"], p["lam"], p["sigTL"], p["sigGs"], p["wp"], p["bl"]],
                                      dtype=np.float64)
        out[f"hbr{k}_cdf"] = np.array([lut[x]["cdf"] for x in range(lut["low"], lut["high"])])
        out[f"hbr{k}_range"] = np.array([lut[x]["range"] for x in range(lut["low"], lut["high"])])
        np.random.seed(200 + k)
        out[f"hbr{k}_out_dn"] = hb.map(data.copy() * span, iso, norm=False)                # DN in, + bl out
    np.savez_compressed(os.path.join(OUT, "realdata.npz"), **out)
    print("wrote realdata.npz", {k: (v.shape, str(v.dtype)) for k, v in out.items() if k.endswith("_out") or "m1_b1" in k})


if __name__ == "__main__":
    main()
