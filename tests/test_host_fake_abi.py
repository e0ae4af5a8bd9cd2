"""Host-side logic of the rows added after round 1's GPU budget was spent, checked on the CPU against a FAKE C ABI: the ctypes
call is intercepted and carried out in NumPy exactly as include/pnnp_b200.h specifies it.  This pins what the Python layer
hands to the library (argument order, dtypes chosen from NumPy's promotion, draw order); the CUDA kernels behind the real ABI
are covered by the opt-in `-m gpu` tests of the same rows."""
import contextlib
import ctypes as C

import numpy as np
import pytest
import torch

import oracle_np as O
from pnnp_b200 import _lib, crops


class _FakeLib:
    """pnnp_wb_gains as the header specifies it, on host memory."""

    def __init__(self):
        self.calls = []

    def pnnp_wb_gains(self, data, n, c, h, w, rgb_gain, kind, gain, stream):
        kind, gain = [int(kind[i]) for i in range(c)], [float(gain[i]) for i in range(c)]
        self.calls.append((n, c, h, w, rgb_gain, kind, gain))
        buf = np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_float)), shape=(n, c, h, w))
        buf *= np.float32(rgb_gain)
        for ch in range(c):
            if kind[ch] == 1:
                buf[:, ch] = buf[:, ch] * np.float32(gain[ch])
            elif kind[ch] == 2:
                buf[:, ch] = (buf[:, ch].astype(np.float64) * np.float64(gain[ch])).astype(np.float32)
        return 0


@pytest.fixture()
def fake_abi(monkeypatch):
    fake = _FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name="tensor": None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda device=None: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    return fake


@pytest.mark.parametrize("tag", ["wb32", "wb64", "wbpy"])
def test_wb_jitter_marshals_what_numpy_would_compute(golden, fake_abi, tag):
    g = golden("wb_jitter")
    wb = [float(v) for v in g[f"{tag}_wb"]] if tag == "wbpy" else g[f"{tag}_wb"]
    gains = tuple(torch.from_numpy(g[f"{tag}_{k}"].copy()) for k in ("rgb", "red", "blue"))
    x = torch.from_numpy(g["base"].copy())
    out = crops.wb_jitter(x, wb, gains)
    assert out is x and out.numpy().tobytes() == g[f"{tag}_out"].tobytes()          # == the unmodified reference's statements
    (n, c, h, w, rgb, kind, gain), = fake_abi.calls
    assert (n, c, h, w) == g["base"].shape and kind[1] == kind[3] == 0
    assert kind[0] == kind[2] == (2 if tag == "wb64" else 1)                        # np.float64 white balance -> float64 product
    assert np.float32(rgb) == g[f"{tag}_rgb"][0]


def test_preprocess_train_draws_like_the_reference_loop(monkeypatch):
    """trainer_SID.py:449-462: one sample_params_max(camera, ratio=None) per crop, in crop order, then the float32 chain with
    the dataset's code / ori / clip, then the clamps of :481-485."""
    from pnnp_b200 import trainer as T
    seen = {}

    def fake_synth(clean, params, noise_code="p", chain=_lib.CHAIN_NUMPY, ori=False, clip=False, **kw):
        seen.update(params=params, code=noise_code, chain=chain, ori=ori, clip=clip, kw=kw)
        return clean + 2.0                                                          # everything above the upper clamp
    monkeypatch.setattr(T, "synthesize_batch", fake_synth)
    tr = T.SID_Trainer.__new__(T.SID_Trainer)
    hr = torch.rand(5, 4, 8, 8) * 1.2 - 0.1
    for cam, clip_cfg in (("IMX686", False), ("SonyA7S2", 2), ("SonyA7S2", 1)):
        cfg = dict(camera_type=cam, noise_code="prq", ori=False, clip=clip_cfg, params=None)
        np.random.seed(31)
        lr, hr2 = tr.preprocess_train(hr.clone(), hr.clone(), cfg)
        np.random.seed(31)
        want = [O.sample_params_max(cam, ratio=None) for _ in range(5)]
        assert len(seen["params"]) == 5
        for a, b in zip(seen["params"], want):
            assert all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in b)
        assert seen["chain"] == _lib.CHAIN_TORCH and seen["code"] == "prq" and seen["clip"] == bool(clip_cfg) and not seen["kw"]
        if clip_cfg:
            assert float(lr.max()) == 1.0 and torch.equal(hr2, hr.clamp(0, 1))
        else:
            assert torch.equal(lr, hr + 2.0) and torch.equal(hr2, hr)
    fixed = O.sample_params_max("SonyA7S2", ratio=150, iso=1600)
    tr.preprocess_train(hr.clone(), hr.clone(), dict(camera_type="SonyA7S2", noise_code="p", ori=False, clip=0, params=fixed))
    assert all(p is fixed for p in seen["params"])                                   # dst.args['params'] short-cuts the draws


@pytest.mark.parametrize("gpu_preprocess", [False, True])
def test_raw_dataset_item_flow_with_wb_jitter(monkeypatch, fake_abi, gpu_preprocess):
    """Raw_Dataset.__getitem__ control flow (syn_datasets.py:296-347) with the device stages replaced by the oracle: crop points
    -> coin -> random_gains -> products -> [per-crop sample_params -> noise from the UNCLIPPED jittered crops] -> clips."""
    import os
    import yaml
    from conftest import ROOT
    from pnnp_b200 import datasets, isp_ops, noise
    cfg = yaml.load(open(os.path.join(ROOT, "runfiles/SonyA7S2/PNNP.yml")), Loader=yaml.FullLoader)["dst_train"]
    cfg.update(H=64, W=96, patch_size=16, crop_per_image=3, lock_wb=False, gpu_preprocess=gpu_preprocess)
    rs = np.random.RandomState(0)
    raw = rs.randint(512, 16384, size=(64, 96)).astype(np.uint16)
    monkeypatch.setattr(datasets.Raw_Dataset, "device", staticmethod(lambda: torch.device("cpu")))
    monkeypatch.setattr(datasets.Raw_Dataset, "synthetic_raw", lambda self, idx, device: torch.from_numpy(raw.view(np.int16).copy()))
    monkeypatch.setattr(isp_ops, "raw2bayer", lambda r, wp, bl, norm, clip: torch.from_numpy(
        O.raw2bayer(r.numpy().view(np.uint16), wp=wp, bl=bl, norm=norm, clip=clip)))
    monkeypatch.setattr(crops, "random_crop", lambda img, hs, ws, aug, patch: torch.from_numpy(
        O.random_crop(img.numpy(), hs, ws, patch, aug)))
    seen = {}

    def fake_synth(clean, params, noise_code="p", ori=False, post_clip=None, **kw):
        seen.update(clean=clean.clone(), params=params, post=post_clip)
        out = clean - 7.0
        return out if post_clip is None else out.clamp(post_clip[0], post_clip[1])
    monkeypatch.setattr(noise, "synthesize_batch", fake_synth)
    ds = datasets.Raw_Dataset(cfg)
    coins = set()
    for seed in range(8):
        seen.clear()
        np.random.seed(seed); torch.manual_seed(seed)
        item = ds[0]
        np.random.seed(seed); torch.manual_seed(seed)
        hs, ws, aug = crops.init_random_crop_point(32, 48, 16, 3, cfg["croptype"])
        hr = O.random_crop(O.raw2bayer(raw, cfg["wp"], cfg["bl"], True, True), hs, ws, 16, aug)
        coin = int(np.random.randint(2))
        coins.add(coin)
        if coin:
            hr = O.wb_jitter(hr, np.ones(4, np.float32), O.random_gains())
            assert hr.max() > 1.0                                                    # the jitter does push crops past 1
        if gpu_preprocess:
            assert not seen and torch.equal(item["ratio"], torch.ones(3))
            assert item["lr"].numpy().tobytes() == hr.clip(-np.inf, 1).tobytes()     # clip == 2: upper clip only
        else:
            params = [O.sample_params("SonyA7S2") for _ in range(3)]
            assert seen["clean"].numpy().tobytes() == hr.tobytes()                   # noise sees the unclipped crops
            assert seen["post"] == (-float("inf"), 1.0)
            assert [float(p["ratio"]) for p in seen["params"]] == [float(p["ratio"]) for p in params]
            assert np.array_equal(item["ratio"].numpy(), np.array([p["ratio"] for p in params], np.float32))
        assert item["hr"].numpy().tobytes() == hr.clip(0, 1).tobytes()
    assert coins == {0, 1}
