import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import pnnp_b200 as P
from pnnp_b200 import train, archs
torch.manual_seed(0)
net = P.UNetSeeInDark({"in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": False}).cuda()
archs.initialize_weights(net)
g = torch.Generator(device="cuda").manual_seed(1)
hr = torch.rand((4, 4, 64, 64), device="cuda", generator=g) ** 2
lr_in = hr + 0.05 * torch.randn((4, 4, 64, 64), device="cuda", generator=g)
ts = train.UNetTrainStep(net, lr=1e-3)
p0 = ts.flat_p.clone()
for i in range(12):
    loss = ts.step(lr_in, hr)
    print(i, float(loss), "dparam max", float((ts.flat_p - p0).abs().max()), "gnorm", float(ts.flat_g.norm()),
          "b10", net.conv10_1.bias.detach().cpu().numpy().round(5), "pred mean", float(ts.scr.bufs['pred'].mean()))
