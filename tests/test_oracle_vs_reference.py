"""Live check of the oracle against the unmodified reference (build container only: skipped on
the GPU box, where /root/reference does not exist)."""
import numpy as np
import pytest
import torch

import oracle_np as O
import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def R():
    return rh.load()


def test_params_live(R):
    P = R.process
    for cam, kw in (("SonyA7S2", {}), ("CRVD", {}), ("SonyA7S2", {"ln_ratio": True})):
        for s in range(8):
            np.random.seed(s); a = P.sample_params(cam, **kw)
            np.random.seed(s); b = O.sample_params(cam, **kw)
            assert all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) and type(a[k]) is type(b[k]) for k in a)
    for cam, kw in (("SonyA7S2", {}), ("SonyA7S2", {"iso": 800}), ("IMX686", {}), ("IMX686", {"iso": 6400}),
                    ("NikonD850", {}), ("CRVD", {"ratio": 7})):
        for s in range(4):
            np.random.seed(s); a = P.sample_params_max(cam, **kw)
            np.random.seed(s); b = O.sample_params_max(cam, **kw)
            assert all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) and type(a[k]) is type(b[k]) for k in a)


def test_noisy_obs_live(R):
    P = R.process
    rng = np.random.RandomState(5)
    y = rng.rand(4, 24, 40).astype(np.float32) ** 2
    for code in ("pgrq", "pg", "prq", "g", "pgrqd", "b"):
        for mk in (lambda: P.sample_params("SonyA7S2"), lambda: P.sample_params_max("SonyA7S2", iso=6400),
                   lambda: P.sample_params_max("IMX686", iso=6400)):
            np.random.seed(2)
            p = mk()
            if "d" in code and not hasattr(p["bias"], "reshape"):
                continue
            np.random.seed(3); a = P.generate_noisy_obs(y, param=p, noise_code=code)
            np.random.seed(3); b, d = O.generate_noisy_obs(y, param=p, noise_code=code, return_draws=True)
            assert a.tobytes() == b.tobytes()
            assert O.noisy_obs_tail_explicit(y, p, code, d).tobytes() == a.tobytes()


def test_networks_live(R):
    arch = dict(name="x", in_nc=4, out_nc=4, nf=8, nframes=1, use_dpsv=False, res=False, cascade=False, add=False,
                lock_wb=False)
    x = torch.rand(1, 4, 32, 32)
    for cls, fn in ((R.archs.UNetSeeInDark, O.unet_forward), (R.archs.ResUnet, O.resunet_forward)):
        net = cls(arch)
        R.archs.initialize_weights(net)
        net.eval()
        with torch.no_grad():
            assert torch.equal(net(x), fn(x, net.state_dict()))


def test_wb_jitter_live(R):
    """random_gains (unprocess.py:60-77) and the statements of syn_datasets.py:314-319 against the oracle, fresh seeds."""
    rs = np.random.RandomState(9)
    base = rs.rand(2, 4, 8, 12).astype(np.float32)
    for s, wb in ((1, np.array([2.0, 1, 1.5, 1], np.float32)), (2, np.array([1.87, 1, 1.61, 1], np.float64)), (3, [2.3, 1.0, 1.4, 1.0])):
        for cam in ("SonyA7S2", "IMX686"):
            np.random.seed(s); torch.manual_seed(s)
            rgb_gain, red_gain, blue_gain = R.syn_datasets.random_gains(cam)
            np.random.seed(s); torch.manual_seed(s)
            gains = O.random_gains(cam)
            assert all(a.numpy().tobytes() == b.tobytes() for a, b in zip((rgb_gain, red_gain, blue_gain), gains))
            hr_crops = base.copy()
            red = wb[0] / red_gain.numpy()
            blue = wb[2] / blue_gain.numpy()
            hr_crops *= rgb_gain.numpy()
            hr_crops[:, 0] = hr_crops[:, 0] * red
            hr_crops[:, 2] = hr_crops[:, 2] * blue
            assert O.wb_jitter(base, wb, gains).tobytes() == hr_crops.tobytes()
