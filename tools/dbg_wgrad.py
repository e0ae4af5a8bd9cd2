"""Debug harness: one wgrad launch per subprocess with PNNP_WG_DBG variants (GPU box only)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def one(ci, co, h, w, n):
    import torch
    import torch.nn.functional as F
    from pnnp_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(1)
    bf = lambda t: t.to(torch.bfloat16).float()
    x = bf(torch.randn((n, ci, h, w), device="cuda", generator=g)); go = bf(torch.randn((n, co, h, w), device="cuda", generator=g))
    wt = torch.zeros((co, ci, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(x, wt, padding=1).backward(go)
    ppad = n * (h + 2) * (w + 2); row = (ppad + 63) // 64 * 64
    gT = torch.empty((co, row), dtype=torch.bfloat16, device="cuda"); xT = torch.empty((ci, row), dtype=torch.bfloat16, device="cuda")
    gon = go.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16); xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    sp = L.stream_ptr(torch.device("cuda"))
    L.check(L.lib().pnnp_transpose_pad(gon.data_ptr(), gT.data_ptr(), n, h, w, co, 0, co, 1, 0, 0, row, sp))
    L.check(L.lib().pnnp_transpose_pad(xn.data_ptr(), xT.data_ptr(), n, h, w, ci, 0, ci, 1, 0, 0, row, sp))
    torch.cuda.synchronize()
    offs = [(dy - 1) * (w + 2) + (dx - 1) for dy in range(3) for dx in range(3)]
    dw = torch.zeros((9, co, ci), device="cuda")
    L.check(L.lib().pnnp_wgrad_tc(gT.data_ptr(), xT.data_ptr(), row, ppad, co, ci, 9, (C.c_int * 9)(*offs), dw.data_ptr(), 0, ci, sp))
    torch.cuda.synchronize()
    got = dw.permute(1, 2, 0).reshape(co, ci, 3, 3)
    print("pipeline_err", L.lib().pnnp_wgrad_pipeline_error(), "rel", ((got - wt.grad).norm() / wt.grad.norm()).item())

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*map(int, sys.argv[1:]))
    else:
        for shape in ["16 32 16 32 1", "32 32 16 32 1", "64 128 16 16 2", "256 256 8 16 1"]:
            for dbg in ["15", "14", "7", "6", "4", "8", "0"]:
                env = dict(os.environ, PNNP_WG_DBG=dbg)
                r = subprocess.run([sys.executable, __file__] + shape.split(), env=env, capture_output=True, text=True, timeout=300)
                tail = (r.stdout + r.stderr).strip().splitlines()[-1:] or [""]
                print(f"shape {shape} dbg {dbg}: rc={r.returncode} {tail[0][:160]}", flush=True)
