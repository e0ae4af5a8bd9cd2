"""Golden vectors for the white-balance jitter of Raw_Dataset.__getitem__ (data_process/syn_datasets.py:313-319) from the
UNMODIFIED reference (build container only):   python oracle/make_golden_wb.py
The dataset class itself needs the SID RAW files, so the reference's statements are executed verbatim here with the
reference's own `random_gains` (data_process/unprocess.py:60-77)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    import torch
    R = rh.load()
    random_gains = R.syn_datasets.random_gains
    out = {}
    rs = np.random.RandomState(11)
    base = (rs.rand(3, 4, 16, 24).astype(np.float32)) ** 2
    for k, (wb, tag) in enumerate(((np.array([2.1, 1.0, 1.6, 1.0], np.float32), "wb32"),
                                   (np.array([1.9321, 1.0, 1.7123, 1.0], np.float64), "wb64"),
                                   ([2.25, 1.0, 1.5, 1.0], "wbpy"))):
        np.random.seed(40 + k)
        torch.manual_seed(40 + k)
        coin = np.random.randint(2)                                   # syn_datasets.py:313 draws this first
        data = {"wb": wb}
        hr_crops = base.copy()
        # ---- verbatim: syn_datasets.py:314-319
        rgb_gain, red_gain, blue_gain = random_gains()
        g = (rgb_gain.numpy().copy(), red_gain.numpy().copy(), blue_gain.numpy().copy())
        red_gain = data['wb'][0] / red_gain.numpy()
        blue_gain = data['wb'][2] / blue_gain.numpy()
        hr_crops *= rgb_gain.numpy()
        hr_crops[:,0] = hr_crops[:,0] * red_gain
        hr_crops[:,2] = hr_crops[:,2] * blue_gain
        # ----
        out[f"{tag}_coin"] = np.array(coin)
        out[f"{tag}_wb"] = np.asarray(wb)
        out[f"{tag}_rgb"], out[f"{tag}_red"], out[f"{tag}_blue"] = g
        out[f"{tag}_out"] = hr_crops
        out[f"{tag}_red_eff_dtype"] = np.array(str(np.asarray(red_gain).dtype))
    np.random.seed(77)
    torch.manual_seed(77)
    g = R.syn_datasets.random_gains(camera_type="IMX686")
    out["imx_rgb"], out["imx_red"], out["imx_blue"] = (t.numpy() for t in g)
    out["base"] = base
    np.savez_compressed(os.path.join(OUT, "wb_jitter.npz"), **out)
    print("wrote wb_jitter.npz", {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
