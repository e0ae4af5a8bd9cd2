"""Where the end-to-end step of the default workload goes (developer tool): CUDA-event times of the stages of
SynthDenoisePipeline.run on 64 crops (one chunk), host time of the step's Python, and the whole step as bench.py times it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pnnp_b200 as P
from pnnp_b200 import _lib
from pnnp_b200.pipeline import SynthDenoisePipeline
from pnnp_b200.isp_ops import raw2bayer
from pnnp_b200.metrics import eval_partial_sums, finish_metrics
n, c, h, w = 64, 4, 512, 512
arch = dict(name="UNetSeeInDark", in_nc=4, out_nc=4, nf=32, nframes=1, use_dpsv=False, res=False, cascade=False, add=False, lock_wb=False)
net = P.UNetSeeInDark(arch).cuda().eval(); P.initialize_weights(net)
np.random.seed(1997)
table = P.ParamTable([P.sample_params("SonyA7S2") for _ in range(n)], "cuda")
raw_host = torch.randint(512, 16383, (n, 2 * h, 2 * w), dtype=torch.int16).pin_memory()
gen = P.PhiloxGenerator(1997)
chunk = int(os.environ.get("PNNP_E2E_CHUNK", "64"))
pipe = SynthDenoisePipeline(net, n, 2 * h, 2 * w, 16383, 512, "pgrq", "cuda", chunk=chunk)
def step():
    sums = pipe.run(raw_host, table=table, generator=gen, next_host=raw_host)
    torch.cuda.current_stream().synchronize()
    return finish_metrics(sums, c, h, w)
for _ in range(3): step()
t0 = time.perf_counter()
for _ in range(10): step()
print(f"whole step (chunk {chunk}): {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
# host-only cost of finish_metrics
sums = pipe.run(raw_host, table=table, generator=gen, next_host=raw_host); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): finish_metrics(sums, c, h, w)
print(f"finish_metrics (host): {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
# stages on the device
d_raw = raw_host.cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
with torch.no_grad():
    for rep in range(3):
        ev[0].record()
        hr = raw2bayer(d_raw, wp=16383, bl=512, norm=True, clip=True)
        ev[1].record()
        lr = P.synthesize_batch(hr, None, "pgrq", post_clip=(-float("inf"), 1.0), table=table, seed_offset=gen.next())
        ev[2].record()
        dn = net(lr)
        ev[3].record()
        s = eval_partial_sums(dn, hr, 1.0, False)
        ev[4].record()
        torch.cuda.synchronize()
names = ["pack (raw2bayer)", "synthesis", "UNet forward", "clamp + PSNR/SSIM sums"]
for i, nm in enumerate(names):
    print(f"{nm:28s} {ev[i].elapsed_time(ev[i + 1]):8.3f} ms")
t0 = time.perf_counter()
for _ in range(10):
    with torch.no_grad():
        hr = raw2bayer(d_raw, wp=16383, bl=512, norm=True, clip=True)
        lr = P.synthesize_batch(hr, None, "pgrq", post_clip=(-float("inf"), 1.0), table=table, seed_offset=gen.next())
        dn = net(lr)
        s = eval_partial_sums(dn, hr, 1.0, False)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue time of one step's launches: {(t1 - t0) / 10 * 1e3:.3f} ms; with execution {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
