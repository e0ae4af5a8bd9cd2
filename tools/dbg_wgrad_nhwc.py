"""Debug harness for csrc/wgrad_nhwc_tc.cu: one launch per subprocess (GPU box only)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def one(mode, ci, co, h, w, n):
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    from pnnp_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(1)
    bf = lambda t: t.to(torch.bfloat16).float()
    x = bf(torch.randn((n, ci, h, w), device="cuda", generator=g))
    if mode == 0:
        go = bf(torch.randn((n, co, h, w), device="cuda", generator=g))
        wt = torch.zeros((co, ci, 3, 3), device="cuda", requires_grad=True)
        F.conv2d(x, wt, padding=1).backward(go)
        taps = 9
    else:
        go = bf(torch.randn((n, co, 2 * h, 2 * w), device="cuda", generator=g))
        wt = torch.zeros((ci, co, 2, 2), device="cuda", requires_grad=True)
        F.conv_transpose2d(x, wt, stride=2).backward(go)
        taps = 4
    gon = go.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16); xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    dw = torch.zeros((taps, ci, co), device="cuda")
    sp = L.stream_ptr(torch.device("cuda"))
    L.check(L.lib().pnnp_wgrad_nhwc(mode, gon.data_ptr(), co, co, xn.data_ptr(), ci, ci, n, h, w, dw.data_ptr(), 0, ci, co, sp))
    torch.cuda.synchronize()
    ref = wt.grad.permute(2, 3, 1, 0).reshape(9, ci, co) if mode == 0 else wt.grad.permute(2, 3, 0, 1).reshape(4, ci, co)
    per_tap = [((dw[t] - ref[t]).norm() / ref[t].norm()).item() for t in range(taps)]
    print("pipeline_err", L.lib().pnnp_wgrad_nhwc_pipeline_error(), "rel", ((dw - ref).norm() / ref.norm()).item(),
          "per tap", " ".join(f"{v:.3f}" for v in per_tap))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*map(int, sys.argv[1:]))
    else:
        for shape in ["0 32 32 16 32 1", "0 16 32 16 32 1", "0 64 64 16 32 1", "0 128 128 16 16 1", "0 64 128 8 16 1", "1 64 32 8 16 1", "1 128 64 8 16 1"]:
            r = subprocess.run([sys.executable, __file__] + shape.split(), capture_output=True, text=True, timeout=300)
            tail = (r.stdout + r.stderr).strip().splitlines()[-1:] or [""]
            print(f"shape {shape}: rc={r.returncode} {tail[0][:220]}", flush=True)
