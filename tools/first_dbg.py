"""Timing of the fused first-layer kernel (developer tool): warp-specialised vs single-role form, and the single-role form with parts
switched off (PNNP_FIRST_DBG bits: 1 no MMA round trip, 2 no global loads, 4 no stores, 8 no im2col)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pnnp_b200 as P
from pnnp_b200 import archs, _lib
x = torch.rand((64, 4, 512, 512), device="cuda")
m = torch.nn.Conv2d(4, 32, 3, padding=1).cuda()
out = torch.empty((64, 512, 512, 32), dtype=torch.bfloat16, device="cuda")
def run(label):
    for _ in range(3): archs._first_conv(x, m, out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): archs._first_conv(x, m, out, 1)
    e1.record(); torch.cuda.synchronize()
    print(f"{label:28s} {e0.elapsed_time(e1)/10*1e3:8.1f} us", flush=True)
os.environ["PNNP_FIRST_WS"] = "1"; run("warp-specialised")
os.environ["PNNP_FIRST_WS"] = "0"; run("single-role")
if "--parts" in sys.argv:
    for dbg in (1, 2, 4, 8, 15):
        os.environ["PNNP_FIRST_DBG"] = str(dbg); run(f"single-role, dbg={dbg}")
print("pipeline error word:", _lib.lib().pnnp_conv_first_pipeline_error())
