"""One metric pass (clamp + PSNR / SSIM partial sums) on 64 crops between cudaProfilerStart / Stop (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnnp_b200.metrics import eval_partial_sums
dn = torch.rand((64, 4, 512, 512), device="cuda")
hr = (dn + 0.02 * torch.randn_like(dn)).clamp(0, 1)
for _ in range(2): eval_partial_sums(dn, hr, 1.0, False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eval_partial_sums(dn, hr, 1.0, False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
