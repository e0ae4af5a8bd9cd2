"""Timing experiments on the specialised noise-synthesis kernel (developer tool; results in DESIGN 4.1 "measured floor"):
PNNP_SYNTH_EXP bits: 1 no Poisson samplers, 2 no Tukey-lambda quantile, 4 no sorting by sampler, 8 no Philox rounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pnnp_b200 as P
from pnnp_b200 import _lib
n, c, h, w = 64, 4, 512, 512
g = torch.Generator(device="cuda").manual_seed(1997)
clean = torch.rand((n, c, h, w), device="cuda", generator=g) ** 2
np.random.seed(1997)
table = P.ParamTable([P.sample_params("SonyA7S2") for _ in range(n)], "cuda")
out = torch.empty_like(clean)
gen = P.PhiloxGenerator(1997)
def run(label):
    f = lambda: P.synthesize_batch(clean, None, "pgrq", _lib.CHAIN_NUMPY, post_clip=(-float("inf"), 1.0), generator=gen, out=out, table=table)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{label:60s} {ms*1e3:8.1f} us   {n*c*h*w*8/ms/1e6:8.1f} GB/s", flush=True)
if os.environ.get("PNNP_SYNTH_ONLY"):
    run("product kernel")
    sys.exit(0)
for e, label in ((0, "product kernel"), (1, "no Poisson samplers"), (2, "no Tukey-lambda quantile"), (3, "no Poisson, no Tukey"),
                 (5, "no Poisson, no sorting"), (7, "no Poisson, no Tukey, no sorting"), (8, "no Philox rounds"),
                 (15, "none of them: loads, rates, queue traffic, f64 tail, stores")):
    if e: os.environ["PNNP_SYNTH_EXP"] = str(e)
    run(f"EXP={e:2d}  {label}")
