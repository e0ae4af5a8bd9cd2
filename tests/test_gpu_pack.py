"""P1/P2 on the GPU through the C ABI vs the oracle and the reference goldens (bit-exact)."""
import numpy as np
import pytest
import torch

import oracle_np as O
import pnnp_b200 as P

pytestmark = pytest.mark.gpu


def _table_raw(n):
    codes = np.arange(n, dtype=np.uint16)
    raw = np.zeros((2, 2 * n), np.uint16)
    raw[:, 0::2] = codes
    raw[:, 1::2] = codes
    return raw


@pytest.mark.parametrize("cam,wp,bl,n", [("sony", 16383, 512, 16384), ("imx686", 1023, 64, 1024)])
def test_exhaustive_sensor_codes(golden, cam, wp, bl, n):
    g = golden("pack")
    raw = _table_raw(n)
    for clip in (0, 1):
        got = P.raw2bayer(raw, wp=wp, bl=bl, norm=True, clip=bool(clip))
        assert got.dtype == np.float32 and got.shape == (4, 1, n)
        for c in range(4):
            assert got[c, 0].tobytes() == g[f"{cam}_clip{clip}"].tobytes()
    packed = P.raw2bayer(raw, wp=wp, bl=bl, norm=True, clip=True)
    rt = P.bayer2raw(packed, wp=wp, bl=bl)
    assert rt.dtype == np.uint16 and np.array_equal(rt, np.clip(raw, bl, wp))


def test_golden_random(golden):
    g = golden("pack")
    raw = g["rand_raw"]
    assert P.raw2bayer(raw, 16383, 512).tobytes() == g["rand_packed"].tobytes()
    assert P.raw2bayer(raw, 16383, 512, clip=True, bias=np.array([1, -2, 3, 0])).tobytes() == g["rand_packed_bias"].tobytes()
    assert P.raw2bayer(raw, 16383, 512, norm=False).tobytes() == g["rand_packed_nonorm"].tobytes()
    assert np.array_equal(P.bayer2raw(torch.from_numpy(g["unpack_in"]), 16383, 512), g["unpack_out"])
    assert np.array_equal(P.bayer2raw(g["unpack_in"][0], 16383, 512), g["unpack_out"])


@pytest.mark.parametrize("H,W", [(2, 2), (6, 10), (16, 64), (18, 50), (128, 272)])
@pytest.mark.parametrize("dtype", ["u16", "f32"])
def test_ragged_shapes_vs_oracle(H, W, dtype):
    rs = np.random.RandomState(H * 1000 + W)
    raw = rs.randint(0, 16384, size=(H, W)).astype(np.uint16)
    if dtype == "f32":
        raw = raw.astype(np.float32) + rs.rand(H, W).astype(np.float32)
    for clip in (False, True):
        for bias in (np.array([0, 0, 0, 0]), np.array([3, -1, 0, 7])):
            want = O.raw2bayer(raw, 16383, 512, True, clip, bias)
            got = P.raw2bayer(raw, 16383, 512, True, clip, bias)
            assert got.tobytes() == want.tobytes()
    x = rs.rand(4, H // 2, W // 2).astype(np.float32) * 1.3 - 0.15
    assert np.array_equal(P.bayer2raw(x, 16383, 512), O.bayer2raw(x, 16383, 512))
    assert np.array_equal(P.bayer2raw(x, 1023, 64), O.bayer2raw(x, 1023, 64))


def test_batched_device_path_full_frame():
    """Full Sony frame size on the device path + round-trip property at full size."""
    rs = np.random.RandomState(7)
    raw = rs.randint(0, 16384, size=(2, 2848, 4256)).astype(np.uint16)
    t = torch.from_numpy(raw.view(np.int16)).cuda()
    packed = P.raw2bayer(t, 16383, 512, clip=True)
    assert packed.shape == (2, 4, 1424, 2128) and packed.is_cuda
    want0 = O.raw2bayer(raw[0], 16383, 512, clip=True)
    assert packed[0].cpu().numpy().tobytes() == want0.tobytes()
    back = P.bayer2raw(packed[1:2], 16383, 512)
    assert np.array_equal(back, np.clip(raw[1], 512, 16383))
