"""Larger HighBitRecovery.map goldens from the UNMODIFIED reference (build container only): 1 x 4 x 64 x 128 dark frames per
camera, `float=True` and `float=False`.  Inputs and uniforms are regenerated from the seeds below by the tests, so only the
reference's outputs are stored:   python oracle/make_golden_hbr_large.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = ((0, "SonyA7S2", "pgrq", 3200), (1, "IMX686", "prq", 6400))
SHAPE = (1, 4, 64, 128)


def inputs(k, sigma, span):
    """The dark frame (normalised float32) and the uniforms of case k; shared with tests/test_realdata_rows.py."""
    data = (np.random.RandomState(40 + k).randn(*SHAPE) * sigma * 1.5 / span).astype(np.float32)
    rand = np.random.RandomState(300 + k).uniform(0, 1, size=SHAPE)
    return data, rand


def main():
    R = rh.load()
    P = R.process
    out = {}
    for k, cam, code, iso in CASES:
        for flt in (True, False):
            np.random.seed(100 + k)
            hb = P.HighBitRecovery(camera_type=cam, noise_code=code, float=flt)
            hb.get_lut([iso], blc_mean=None)
            lut = hb.lut[iso]
            span = lut["param"]["wp"] - lut["param"]["bl"]
            data, rand = inputs(k, lut["sigma"], span)
            np.random.seed(300 + k)                     # RandomState(300 + k).uniform == the global stream after seed(300 + k)
            res = hb.map(data.copy(), iso, norm=True)
            assert res.dtype == np.float32
            out[f"hbrL{k}_{'float' if flt else 'int'}"] = res
    np.savez_compressed(os.path.join(OUT, "hbr_large.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
