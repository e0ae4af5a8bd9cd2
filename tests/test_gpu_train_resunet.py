"""ResUnet synthetic-pair training step (train_resunet.ResUnetTrainStep) against fp32 autograd through the oracle's functional
ResUnet (oracle_np.resunet_forward, pinned to the reference module) and against the reference's L1 + Adam loop."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle_np as O
import pnnp_b200 as P
from pnnp_b200 import _lib, train_resunet

pytestmark = pytest.mark.gpu
L = _lib
ARCH = dict(name="ResUnet", in_nc=4, out_nc=4, nf=32, nframes=1, use_dpsv=False, res=False, cascade=False, add=False, lock_wb=False)


def _make(seed=3, shape=(2, 4, 64, 96), std=None):
    torch.manual_seed(seed)
    net = P.ResUnet(ARCH).cuda()
    P.initialize_weights(net)
    if std is not None:                                   # O(1) activations: exercises the ReLU masks harder than sigma = 0.02
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() == 4:
                    p.mul_(std / 0.02 / (p.shape[1] * p.shape[2] * p.shape[3]) ** 0.5 * 0.02 * 50)
    g = torch.Generator(device="cuda").manual_seed(seed)
    hr = torch.rand(shape, device="cuda", generator=g) ** 2
    lr = (hr + 0.05 * torch.randn(shape, device="cuda", generator=g)).contiguous()
    return net, lr, hr


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _cos(a, b):
    return F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def test_resunet_backward_matches_fp32_autograd():
    """Every parameter gradient of one step vs autograd through the fp32 functional network, both driven by the SAME d loss / d pred.
    bf16 storage re-routes a few ReLU masks, hence relative L2 < 15 % and cosine > 0.99 per parameter (the bounds of the UNet test)."""
    net, lr_in, hr = _make()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ts = train_resunet.ResUnetTrainStep(net)
    pred, saved = ts.forward(lr_in)
    gp = torch.empty_like(pred)
    L.check(L.lib().pnnp_l1_loss(pred.data_ptr(), hr.data_ptr(), gp.data_ptr(), pred.numel(), ts.loss_sum.data_ptr(),
                                 L.stream_ptr(pred.device)), "l1")
    ts.backward(gp, saved)
    torch.cuda.synchronize()
    assert L.lib().pnnp_conv_pipeline_error() == 0 and L.lib().pnnp_wgrad_nhwc_pipeline_error() == 0
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    pred_ref = O.resunet_forward(lr_in, params)
    pred_ref.backward(gp)
    assert (pred - pred_ref).abs().max().item() < 2e-2 * max(1.0, pred_ref.abs().max().item())
    stats = {name: (_rel(ts._grad_view(name), params[name].grad), _cos(ts._grad_view(name), params[name].grad)) for name in sd}
    worst_rel = max(stats.items(), key=lambda kv: kv[1][0])
    worst_cos = min(stats.items(), key=lambda kv: kv[1][1])
    print("worst rel", worst_rel, "worst cos", worst_cos)
    bad = {k: v for k, v in stats.items() if not (v[1] > 0.99 and v[0] < 0.15)}
    assert not bad, bad


def test_resunet_training_follows_the_reference_loop():
    """Six Adam steps on fixed crops: the loss curve tracks the fp32 reference loop (functional ResUnet + autograd + torch Adam) to
    2 %, the parameters stay close, and the loss falls."""
    net, lr_in, hr = _make(seed=5)
    with torch.no_grad():
        net.conv10.bias.fill_(0.05)                      # the clamp in the loss kills negative outputs' gradient: start all channels alive
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ts = train_resunet.ResUnetTrainStep(net, lr=1e-3)
    losses = [float(ts.step(lr_in, hr)) for _ in range(6)]
    torch.cuda.synchronize()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    ref = []
    for _ in range(6):
        opt.zero_grad()
        loss = F.l1_loss(O.resunet_forward(lr_in, params).clamp(0, 1), hr)
        loss.backward()
        opt.step()
        ref.append(loss.item())
    print("losses", losses, "reference", ref)
    assert all(abs(a - b) < 0.02 * max(b, 1e-3) for a, b in zip(losses, ref)), (losses, ref)
    assert losses[-1] < losses[0]
    drift = max(_rel(p.detach(), params[k]) for k, p in net.named_parameters() if p.dim() == 4)
    assert drift < 0.05, drift


def test_sid_trainer_trains_a_resunet(tmp_path, monkeypatch):
    """`--mode train` with arch.name = ResUnet (trainer_SID.py:17 resolves the class by name): the loop runs, the loss moves, a
    checkpoint with the reference's state_dict keys is written."""
    import os
    import re
    import yaml
    from conftest import ROOT
    from pnnp_b200 import trainer as T
    monkeypatch.chdir(tmp_path)
    cfg = yaml.load(open(os.path.join(ROOT, "runfiles/SonyA7S2/PNNP.yml")), Loader=yaml.FullLoader)
    for k in ("dst", "dst_train", "dst_eval", "dst_test"):
        cfg[k]["H"], cfg[k]["W"], cfg[k]["synthetic_frames"] = 256, 384, 1
        if "iso_list" in cfg[k]:
            cfg[k]["iso_list"] = cfg[k]["iso_list"][:1]
    cfg["arch"]["name"] = "ResUnet"
    cfg["model_name"] += "_ResUnet"
    cfg["fast_ckpt"], cfg["checkpoint"] = str(tmp_path / "ckpt"), str(tmp_path / "saved")
    cfg["dst_train"].update(H=256, W=384, patch_size=64, crop_per_image=4, synthetic_frames=4)
    cfg["hyper"].update(stop_epoch=4, save_freq=2, plot_freq=4, batch_size=2, learning_rate=1e-3, lr_scheduler="MultiStep", step_size=100)
    runfile = tmp_path / "run.yml"
    runfile.write_text(yaml.dump(cfg))
    np.random.seed(5)
    torch.manual_seed(5)
    tr = T.SID_Trainer(["-f", str(runfile), "--mode", "train"])
    tr.net.conv10.bias.data.fill_(0.05)
    step = tr.train()
    text = open(tmp_path / "logs" / f"log_{cfg['model_name']}.log").read()
    l1 = [float(x) for x in re.findall(r"L1=(\d+\.\d+)", text)]
    assert len(l1) == 4 and step.t == 8 and l1[-1] < l1[0], l1
    sd = torch.load(os.path.join(cfg["fast_ckpt"], f"{cfg['model_name']}_last_model.pth"))
    assert "conv1.block.0.conv.conv.weight" in sd and "conv10.weight" in sd
