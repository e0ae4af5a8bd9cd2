"""Golden scalars for the training-step row (T1), from the UNMODIFIED reference modules (build container only):
archs.UNetSeeInDark + archs.initialize_weights + losses.Unet_Loss + torch.optim.Adam(lr=1e-4), the loop body of
trainer_SID.py:93-101, three steps on a seeded 2x4x32x48 pair.     python oracle/make_golden_train.py"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
ARCH = dict(name="UNetSeeInDark", in_nc=4, out_nc=4, nf=32, nframes=1, use_dpsv=False, res=False, cascade=False, add=False, lock_wb=False)


def main():
    R = rh.load()
    torch.set_num_threads(4)
    torch.manual_seed(1997)
    net = R.archs.UNetSeeInDark(ARCH)
    R.archs.initialize_weights(net)
    net.conv10_1.bias.data.fill_(0.05)                    # all four outputs inside the clamp (see tests/test_gpu_trainer.py)
    g = torch.Generator().manual_seed(7)
    hr = torch.rand((2, 4, 32, 48), generator=g) ** 2
    lr = hr + 0.05 * torch.randn((2, 4, 32, 48), generator=g)
    loss_fn = R.losses.Unet_Loss()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    out = {"losses": [], "grad_abs_sum_step0": {}, "param_abs_sum_after": {}}
    net.train()
    for step in range(3):
        opt.zero_grad()
        pred = net(lr)
        loss = loss_fn(pred.clamp(0, 1), hr)
        loss.backward()
        if step == 0:
            out["pred0_abs_sum"] = float(pred.detach().abs().sum())
            for k, p in net.named_parameters():
                out["grad_abs_sum_step0"][k] = float(p.grad.abs().sum())
        opt.step()
        out["losses"].append(float(loss))
    for k, p in net.named_parameters():
        out["param_abs_sum_after"][k] = float(p.detach().double().abs().sum())
    out["versions"] = rh.versions()
    with open(os.path.join(OUT, "train_step.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("losses", out["losses"])


if __name__ == "__main__":
    main()
