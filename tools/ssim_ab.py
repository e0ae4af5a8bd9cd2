"""Same-box comparison of two builds of the metric pass: tools/ssim_ab.py <library.so> times eval_partial_sums on 64 crops of 4x512x512
with that library loaded instead of the in-tree one (developer tool; the product always loads pnnp_b200/libpnnp_b200.so)."""
import subprocess
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 2 and sys.argv[1] == "child":
    import torch
    import pnnp_b200._lib as L
    if sys.argv[2] != "-":
        L.LIB_PATH = os.path.abspath(sys.argv[2])
    from pnnp_b200.metrics import eval_partial_sums
    g = torch.Generator(device="cuda").manual_seed(0)
    hr = torch.rand((64, 4, 512, 512), device="cuda", generator=g)
    dn = (hr + 0.02 * torch.randn(hr.shape, device="cuda", generator=g)).contiguous()
    for _ in range(3):
        s = eval_partial_sums(dn, hr, 1.0, False)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eval_partial_sums(dn, hr, 1.0, False)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
    print(f"{best:8.1f} us per 64 crops   sums[0] = {s[0].tolist()}")
else:
    for rnd in range(2):
        for lib in ["-"] + sys.argv[1:]:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "child", lib], capture_output=True, text=True)
            print(f"{lib:40s} {out.stdout.strip() or out.stderr.strip()[-400:]}", flush=True)
