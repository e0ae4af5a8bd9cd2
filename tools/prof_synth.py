"""One launch of the specialised noise kernel on BASELINE configs[1] (for ncu; developer tool)."""
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
for _ in range(3):
    P.synthesize_batch(clean, None, "pgrq", _lib.CHAIN_NUMPY, post_clip=(-float("inf"), 1.0), generator=gen, out=out, table=table)
torch.cuda.synchronize()
