"""Where a full-resolution x-mode layer's time goes (developer tool): the generic instantiation of conv_gemm_tc_kernel with parts
switched off through PNNP_CONV_DBG (1 skip stores, 2 skip MMAs, 4 skip A loads, 8 skip the epilogue incl. its TMEM loads) next to the
specialised product kernel.  64 crops of 512x512, 32 -> 32 channels + fused pool (conv1_2) and 64 -> 32 two sources (conv9_1)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pnnp_b200 as P
from pnnp_b200 import archs, _lib
n, h, w = 64, 512, 512
g = torch.Generator(device="cuda").manual_seed(1)
def layer(cin, cout, two):
    ct = cin * (2 if two else 1)
    m = type("M", (), {})()
    m.weight, m.bias = torch.randn((cout, ct, 3, 3), device="cuda", generator=g) / (3 * ct ** 0.5), None
    return archs._PackedLayer(m, "conv3x").get("cuda")[0], torch.zeros(cout, device="cuda")
x = torch.randn((n, h, w, 32), device="cuda", generator=g).to(torch.bfloat16)
x2 = torch.randn((n, h, w, 32), device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty((n, h, w, 32), dtype=torch.bfloat16, device="cuda")
pooled = torch.empty((n, h // 2, w // 2, 32), dtype=torch.bfloat16, device="cuda")
w1, b1 = layer(32, 32, False)
w2, b2 = layer(32, 32, True)
def run(label, fn):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:58s} {e0.elapsed_time(e1)/5*1e3:8.1f} us", flush=True)
f1 = lambda: archs._conv(_lib.CONV3X, x, w1, b1, out, 32, _lib.ACT_LEAKY, pool_out=pooled)
f2 = lambda: archs._conv(_lib.CONV3X, x, w2, b2, out, 32, _lib.ACT_LEAKY, x1=x2)
for name, f in (("conv1_2-like (32->32 + pool)", f1), ("conv9_1-like (32+32->32)", f2)):
    os.environ.pop("PNNP_CONV_DBG", None); run(f"{name}: product kernel", f)
    os.environ["PNNP_CONV_NOSPEC"] = "1"
    for dbg, what in ((0, "generic epilogue"), (1, "no stores"), (8, "no epilogue (no TMEM loads)"), (2, "no MMAs"), (4, "no A loads"),
                      (6, "no MMAs, no A loads"), (9, "no epilogue, no stores"), (14, "no epilogue, MMAs, loads: hand-shakes only")):
        os.environ["PNNP_CONV_DBG"] = str(dbg) if dbg else ""
        if not dbg: os.environ.pop("PNNP_CONV_DBG")
        run(f"  generic, dbg={dbg:2d}: {what}", f)
    os.environ.pop("PNNP_CONV_NOSPEC", None); os.environ.pop("PNNP_CONV_DBG", None)
print("pipeline error word:", _lib.lib().pnnp_conv_pipeline_error())
