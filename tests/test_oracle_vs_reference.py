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
