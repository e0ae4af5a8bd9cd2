"""One UNetSeeInDark forward on 64 crops of 4x512x512 between cudaProfilerStart / Stop (for `ncu --profile-from-start off`;
developer tool).  --resunet profiles the ResUnet."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pnnp_b200 as P
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
arch = dict(name="UNetSeeInDark", in_nc=4, out_nc=4, nf=32, nframes=1, use_dpsv=False, res=False, cascade=False, add=False, lock_wb=False)
net = (P.ResUnet if "--resunet" in sys.argv else P.UNetSeeInDark)(arch).cuda().eval(); P.initialize_weights(net)
x = torch.rand((n, 4, 512, 512), device="cuda")
with torch.no_grad():
    for _ in range(2): net(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
