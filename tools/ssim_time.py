import os, sys, torch
sys.path.insert(0, "/root/repo")
from pnnp_b200.metrics import eval_partial_sums
g = torch.Generator(device="cuda").manual_seed(0)
hr = torch.rand((16, 4, 512, 512), device="cuda", generator=g)
dn = (hr + 0.02 * torch.randn(hr.shape, device="cuda", generator=g)).contiguous()
for v in ("1", "0"):
    os.environ["PNNP_SSIM_V2"] = v
    for _ in range(3): eval_partial_sums(dn, hr, 1.0, False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eval_partial_sums(dn, hr, 1.0, False)
    e1.record(); torch.cuda.synchronize()
    print("PNNP_SSIM_V2=" + v, e0.elapsed_time(e1) / 20 * 1e3, "us per 16 crops")
