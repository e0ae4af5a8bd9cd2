"""D2 on the device vs the oracle (numpy rot90 / flip / slicing): bit-exact gather."""
import numpy as np
import pytest
import torch

import oracle_np as O
from pnnp_b200 import crops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w,patch,n", [(64, 96, 32, 8), (1424, 2128, 512, 8), (40, 40, 40, 3), (72, 100, 16, 70)])
def test_random_crop_and_all_eight_aug_modes(h, w, patch, n):
    rs = np.random.RandomState(h + w)
    img = rs.rand(4, h, w).astype(np.float32)
    np.random.seed(5)
    hs, ws, aug = crops.init_random_crop_point(h, w, patch, n)
    aug = np.arange(n) % 8                                  # make sure every mode is exercised
    want = O.random_crop(img, hs, ws, patch, aug)
    got = crops.random_crop(torch.from_numpy(img).cuda(), hs, ws, aug, patch).cpu().numpy()
    assert got.tobytes() == want.tobytes()


def test_crop_points_follow_reference_draw_order():
    np.random.seed(3)
    hs, ws, aug = crops.init_random_crop_point(1424, 2128, 512, 8)
    np.random.seed(3)
    aug2 = np.random.randint(8, size=8)
    pts = [(np.random.randint(0, 1424 - 512 + 1), np.random.randint(0, 2128 - 512 + 1)) for _ in range(8)]
    assert list(aug) == list(aug2) and list(zip(hs, ws)) == pts
    np.random.seed(4)
    hs, ws, _ = crops.init_random_crop_point(1424, 2128, 512, 8, mode='non-overlapped')
    assert len(hs) == (1424 // 512) * (2128 // 512)


def test_data_aug_matches_numpy():
    rs = np.random.RandomState(0)
    x = rs.rand(4, 24, 24).astype(np.float32)
    for mode in range(8):
        got = crops.data_aug(torch.from_numpy(x).cuda(), mode).cpu().numpy()
        assert np.array_equal(got, np.ascontiguousarray(O.data_aug(x, mode)))


def test_raw_dataset_item_on_device():
    """Raw_Dataset.__getitem__ equivalent: shapes, clips and the reference's draw order for crops + params."""
    import yaml, os
    from conftest import ROOT
    from pnnp_b200.datasets import Raw_Dataset
    cfg = yaml.load(open(os.path.join(ROOT, "runfiles/SonyA7S2/PNNP.yml")), Loader=yaml.FullLoader)["dst_train"]
    cfg.update(H=512, W=768, patch_size=128, crop_per_image=8)
    ds = Raw_Dataset(cfg)
    np.random.seed(11)
    item = ds[0]
    assert item["lr"].shape == (8, 4, 128, 128) and item["hr"].shape == (8, 4, 128, 128) and item["ratio"].shape == (8,)
    assert item["lr"].is_cuda and float(item["hr"].min()) >= 0 and float(item["hr"].max()) <= 1
    assert float(item["lr"].max()) <= 1.0 and float(item["lr"].min()) < 0     # clip == 2: upper clip only
    assert 100 <= float(item["ratio"].min()) and float(item["ratio"].max()) <= 300
    # the clean crops are exactly the oracle's crops of the oracle's packed frame
    raw = ds.synthetic_raw(0, item["lr"].device).cpu().numpy().view(np.uint16)
    packed = O.raw2bayer(raw, cfg["wp"], cfg["bl"], True, True)
    np.random.seed(11)
    hs, ws, aug = crops.init_random_crop_point(256, 384, 128, 8, cfg["croptype"])
    assert item["hr"].cpu().numpy().tobytes() == O.random_crop(packed, hs, ws, 128, aug).tobytes()


def test_eval_crop_and_merge_match_reference_goldens(golden):
    """Overlapped tiling for tile-wise inference (syn_datasets.py:109-159): bit-exact vs the reference's own output."""
    g = golden("tiling")
    k = 0
    while f"case{k}_geom" in g:
        c, h, w, patch, base = (int(v) for v in g[f"case{k}_geom"])
        x = torch.from_numpy(g[f"case{k}_x"]).cuda()
        tiles = crops.eval_crop(x, patch, base)
        assert tiles.cpu().numpy().tobytes() == g[f"case{k}_tiles"].tobytes()
        marked = tiles + torch.arange(tiles.shape[0], dtype=torch.float32, device="cuda").view(-1, 1, 1, 1)
        assert crops.eval_merge(marked, h, w, base).cpu().numpy().tobytes() == g[f"case{k}_merged"].tobytes()
        k += 1
    assert k == 5


def test_eval_tiling_round_trip_at_frame_size():
    """4x1424x2128 Sony frame, 512 tiles with 64 overlap: crop -> merge is the identity; the tile grid is the reference's."""
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((1, 4, 1424, 2128), device="cuda", generator=gen)
    tiles = crops.eval_crop(x, 512, 64)
    assert tiles.shape == (4 * 5, 4, 512, 512) and crops.tile_grid(1424, 2128, 512, 64) == (4, 5)
    assert torch.equal(crops.eval_merge(tiles, 1424, 2128, 64), x)
    assert torch.equal(tiles[0, :, 32:, 32:], x[0, :, :480, :480])            # first tile = reflect-padded top-left corner
    assert torch.equal(tiles[0, :, 0, 32:], x[0, :, 32, :480])                # reflect (no edge repeat): padded row -32 = row 32
