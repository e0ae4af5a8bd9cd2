"""Eval entry points (trainer_SID.py / trainer_LRID.py equivalents) end to end on small synthetic frames."""
import os
import re

import pytest
import yaml

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _small_runfile(tmp_path, src, H, W, frames, name=None):
    cfg = yaml.load(open(os.path.join(ROOT, src)), Loader=yaml.FullLoader)
    for k in ("dst", "dst_train", "dst_eval", "dst_test"):
        cfg[k]["H"], cfg[k]["W"] = H, W
        cfg[k]["synthetic_frames"] = frames
        if "iso_list" in cfg[k]:
            cfg[k]["iso_list"] = cfg[k]["iso_list"][:1]
    if name:
        cfg["arch"]["name"] = name
        cfg["model_name"] += "_" + name
    cfg["fast_ckpt"] = str(tmp_path / "ckpt")
    cfg["checkpoint"] = str(tmp_path / "saved")
    p = tmp_path / "run.yml"
    p.write_text(yaml.dump(cfg))
    return str(p), cfg


@pytest.mark.parametrize("arch", ["UNetSeeInDark", "ResUnet"])
def test_sid_evaltest_entry_point(tmp_path, monkeypatch, arch):
    from pnnp_b200 import trainer as T
    monkeypatch.chdir(tmp_path)
    runfile, cfg = _small_runfile(tmp_path, "runfiles/SonyA7S2/PNNP.yml", 256, 384, 2, arch)
    res = T.main_sid(["-f", runfile, "--mode", "evaltest"])
    assert set(res) == {"eval_x100", "eval_x200", "test_x100", "test_x250", "test_x300"}
    for v in res.values():
        assert v["frames"] == 2 and 0 < v["PSNR"] < 80 and -1 <= v["SSIM"] <= 1
    text = open(tmp_path / "logs" / f"log_{cfg['model_name']}.log").read()
    assert "ELD Datasets: Dgain=100" in text and "SID Datasets: Dgain=300" in text
    assert re.search(r">>  Epoch -1: PSNR=\d+\.\d\d\npsnrs_lr=\d+\.\d\d, psnrs_dn=\d+\.\d\d\nssims_lr=-?\d\.\d{4}, ssims_dn=-?\d\.\d{4}", text)
    assert os.path.exists(tmp_path / "metrics" / f"{cfg['model_name']}_metrics.pkl")
    assert os.path.exists(os.path.join(cfg["fast_ckpt"], f"{cfg['model_name']}_last_model.pth"))


def test_lrid_eval_entry_point_takes_reflect_pad_branch(tmp_path, monkeypatch):
    from pnnp_b200 import trainer as T
    monkeypatch.chdir(tmp_path)
    runfile, cfg = _small_runfile(tmp_path, "runfiles/IMX686/PNNP.yml", 80, 112, 2)     # 40 x 56 packed: 56 % 16 != 0
    res = T.main_lrid(["-f", runfile, "--mode", "eval"])
    assert set(res) == {f"eval_x{r}" for r in (1, 2, 4, 8, 16)}
    assert all(v["frames"] == 2 and v["PSNR"] > 0 for v in res.values())


def test_sid_train_mode_runs_the_reference_loop_and_learns(tmp_path, monkeypatch):
    """`--mode train` (trainer_SID.py:74-180): Raw_Dataset items built on the device -> explicit training step; the L1 loss
    falls over ten epochs, a checkpoint with the reference's state_dict keys is written, then the eval sweeps run."""
    from pnnp_b200 import trainer as T
    monkeypatch.chdir(tmp_path)
    runfile, cfg = _small_runfile(tmp_path, "runfiles/SonyA7S2/PNNP.yml", 256, 384, 1)
    cfg["dst_train"].update(H=256, W=384, patch_size=64, crop_per_image=4, synthetic_frames=4)
    cfg["hyper"].update(stop_epoch=10, save_freq=5, plot_freq=10, batch_size=2, learning_rate=1e-3, lr_scheduler="MultiStep",
                        step_size=100)
    open(runfile, "w").write(yaml.dump(cfg))
    import numpy as np
    import torch
    np.random.seed(5)                  # crop positions / noise parameters come from NumPy's global state, as in the reference
    torch.manual_seed(5)
    tr = T.SID_Trainer(["-f", runfile, "--mode", "train"])
    # The reference's loss clamps BEFORE the L1 (losses/base_loss.py:92-103): an output channel whose N(0, 0.02) bias starts
    # negative predicts < 0 everywhere and gets zero gradient.  Start all four channels alive so the loss can move.
    tr.net.conv10_1.bias.data.fill_(0.05)
    step = tr.train()
    text = open(tmp_path / "logs" / f"log_{cfg['model_name']}.log").read()
    l1 = [float(x) for x in re.findall(r"L1=(\d+\.\d+)", text)]
    # 20 Adam steps at lr 1e-3.  The same loop in fp32 on the CPU (oracle UNet + autograd + Adam on statistically identical
    # items) falls from 0.297 to the constant-predictor plateau ~0.25 (-15 %) by epoch 8 and is at -10 % after 6 epochs; the r01
    # GPU run measured -7 % after 6.  Per-epoch means still scatter (two batches each), hence the minimum over the second half.
    assert len(l1) == 10 and min(l1[5:]) < 0.97 * l1[0], l1
    assert step.t == 10 * 2                                           # 4 items / batch 2 = 2 steps per epoch
    assert "Epoch 10: PSNR=" in text                                  # the fast eval at plot_freq
    sd = torch.load(os.path.join(cfg["fast_ckpt"], f"{cfg['model_name']}_last_model.pth"))
    assert "conv1_1.weight" in sd and "upv6.weight" in sd and len(sd) == 46


def test_predict_tiles_a_full_frame_like_the_reference(tmp_path, monkeypatch):
    """trainer_SID.py:345-360: raw2bayer(raw + bl) -> eval_crop -> net per tile -> eval_merge -> <name>.npy."""
    import numpy as np
    import torch
    import pnnp_b200 as P
    from pnnp_b200 import trainer as T
    monkeypatch.chdir(tmp_path)
    runfile, cfg = _small_runfile(tmp_path, "runfiles/SonyA7S2/PNNP.yml", 256, 384, 1)
    for k in ("dst", "dst_train", "dst_eval", "dst_test"):
        cfg[k]["patch_size"] = 96                      # l = 96 - 64 = 32: a 5 x 7 grid of overlapped tiles on the 128 x 192 frame
    open(runfile, "w").write(yaml.dump(cfg))
    tr = T.SID_Trainer(["-f", runfile, "--mode", "evaltest"])
    raw = np.random.RandomState(0).randint(0, 900, size=(256, 384)).astype(np.float32)
    want_in = P.raw2bayer(torch.from_numpy(raw + cfg["dst_eval"]["bl"]).cuda())            # reference defaults wp=1023, bl=64
    net = tr.net
    tr.net = type("Identity", (), {"eval": lambda self: self, "__call__": lambda self, x: x})()
    out = tr.predict(raw, name=str(tmp_path / "ident"))
    assert out.shape == (4, 128, 192) and np.array_equal(out, want_in.reshape(4, 128, 192).cpu().numpy())   # crop -> merge = identity
    assert np.array_equal(np.load(tmp_path / "ident.npy"), out)
    tr.net = net
    dn = tr.predict(raw, name=str(tmp_path / "dn"))
    tiles = tr.dst_eval.eval_crop(want_in.reshape(1, 4, 128, 192))
    with torch.no_grad():
        ref = tr.dst_eval.eval_merge(net.eval()(tiles))[0].cpu().numpy()
    assert np.array_equal(dn, ref) and np.isfinite(dn).all()
