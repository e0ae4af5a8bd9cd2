"""Per-layer CUDA-event timing of the tcgen05 UNet forward (developer tool, not a bench value)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pnnp_b200 as P
from pnnp_b200 import archs, _lib

args = [a for a in sys.argv[1:] if not a.startswith("--")]
shape = tuple(int(v) for v in (args[0] if args else "1,4,1424,2128").split(","))
RES = "--resunet" in sys.argv
arch = dict(name="UNetSeeInDark", in_nc=4, out_nc=4, nf=32, nframes=1, use_dpsv=False, res=False, cascade=False, add=False, lock_wb=False)
net = (P.ResUnet if RES else P.UNetSeeInDark)(arch).cuda().eval(); P.initialize_weights(net)
x = torch.rand(shape, device="cuda")
records = []
orig_conv, orig_pool, orig_in, orig_first = archs._conv, archs._pool, archs._to_nhwc16, archs._first_conv
def timed(fn, label):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        records.append((label(a, k), e0, e1)); return r
    return w
def conv_label(a, k):
    mode, x0, wt = a[0], a[1], a[2]
    n, h, w, c0 = x0.shape
    c1 = 0 if k.get("x1") is None else k["x1"].shape[3]
    cout = a[5]; taps = {0: 9, 1: 1, 2: 4, 3: 9, 4: 9, 6: 9}[mode]
    flops = 2.0 * n * h * w * (c0 + c1) * cout * taps / (4 if mode == 3 else 1)
    extra = ("+resid" if k.get("resid") is not None else "") + ("+pool" if k.get("pool_out") is not None else "") + ("+head" if k.get("head") is not None else "")
    return ({0: "conv", 1: "1x1", 2: "convT", 3: "convS2", 4: "convX", 6: "convB"}[mode], f"{c0+c1}->{cout} @{h}x{w}{extra}", flops)
with torch.no_grad():
    for _ in range(3): net(x)
    torch.cuda.synchronize()
    archs._conv = timed(orig_conv, conv_label)
    archs._pool = timed(orig_pool, lambda a, k: ("pool", f"{a[0].shape[3]} @{a[0].shape[1]}x{a[0].shape[2]}", 0.0))
    archs._to_nhwc16 = timed(orig_in, lambda a, k: ("in", "", 0.0))
    archs._first_conv = timed(orig_first, lambda a, k: ("first", f"{a[0].shape[1]}->{a[1].weight.shape[0]} @{a[0].shape[2]}x{a[0].shape[3]} (fused in+conv)",
                                                        2.0 * a[0].shape[0] * a[0].shape[2] * a[0].shape[3] * a[0].shape[1] * a[1].weight.shape[0] * 9))
    net(x); torch.cuda.synchronize()
    tot = 0.0
    for (kind, desc, fl), e0, e1 in records:
        ms = e0.elapsed_time(e1); tot += ms
        print(f"{kind:6s} {desc:44s} {ms*1e3:9.1f} us  {fl/ms/1e9 if ms>0 else 0:8.1f} TFLOP/s")
    archs._conv, archs._pool, archs._to_nhwc16, archs._first_conv = orig_conv, orig_pool, orig_in, orig_first
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    px = shape[0] * shape[1] * shape[2] * shape[3]
    print(f"sum of layers {tot:.3f} ms; whole forward {ms:.3f} ms; {px/1e6/ms*1e3:.0f} MP/s; {(119424 if RES else 92288)*px/ms/1e9:.1f} TFLOP/s effective")
