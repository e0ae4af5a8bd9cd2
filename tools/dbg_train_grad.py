"""Per-parameter gradient error of UNetTrainStep vs fp32 autograd through the bf16-storage network (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import test_gpu_train as T
from pnnp_b200 import train, _lib as L

for std in (1.4,):
    net, lr_in, hr = T._make(std=std)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ts = train.UNetTrainStep(net)
    pred, saved = ts.forward(lr_in)
    gp = torch.empty_like(pred)
    L.check(L.lib().pnnp_l1_loss(pred.data_ptr(), hr.data_ptr(), gp.data_ptr(), pred.numel(), ts.loss_sum.data_ptr(), T._sp()))
    ts.backward(gp, saved)
    p16 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    # keep intermediate activations of the reference to compare forward values too
    pr = T._unet_forward_bf16_storage(lr_in, p16)
    pr.backward(gp)
    print("pred max abs diff", (pred - pr).abs().max().item(), "pred absmax", pr.abs().max().item())
    order = ["conv10_1", "conv9_2", "conv9_1", "upv9", "conv8_2", "conv8_1", "upv8", "conv7_2", "conv7_1", "upv7", "conv6_2", "conv6_1",
             "upv6", "conv5_2", "conv5_1", "conv4_2", "conv4_1", "conv3_2", "conv3_1", "conv2_2", "conv2_1", "conv1_2", "conv1_1"]
    for n in order:
        for suf in (".weight", ".bias"):
            k = n + suf
            print(f"{k:18s} rel {T._rel(ts._grad_view(k), p16[k].grad):.4f} cos {T._cos(ts._grad_view(k), p16[k].grad):.5f}  |g| {p16[k].grad.norm().item():.3e}")

# ---- intermediate activation gradients: ours (scratch buffers after backward) vs autograd through the bf16-storage network
import torch.nn.functional as F
def fwd_keep(x, sd):
    ste = T._ste
    Z = {}
    def cv(t, n):
        z = F.conv2d(t, ste(sd[n + ".weight"]), sd[n + ".bias"], padding=sd[n + ".weight"].shape[-1] // 2)
        z.retain_grad(); Z["z_" + n] = z
        return z
    act = lambda t: ste(F.leaky_relu(t, 0.2))
    def up(t, n):
        u = ste(F.conv_transpose2d(t, ste(sd[n + ".weight"]), sd[n + ".bias"], stride=2))
        u.retain_grad(); Z["u_" + n] = u
        return u
    c, cur = {}, ste(x)
    for i in range(1, 6):
        c[i] = act(cv(act(cv(cur, f"conv{i}_1")), f"conv{i}_2"))
        cur = F.max_pool2d(c[i], 2) if i < 5 else c[i]
    for i in range(6, 10):
        cur = act(cv(act(cv(torch.cat([up(cur, f"upv{i}"), c[10 - i]], 1), f"conv{i}_1")), f"conv{i}_2"))
    return cv(cur, "conv10_1"), Z

p16 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
pr, Z = fwd_keep(lr_in, p16)
pr.backward(gp)
B = ts.scr.bufs
def cmp(tag, ours_nhwc, ref_nchw):
    o = ours_nhwc.float().permute(0, 3, 1, 2)
    print(f"{tag:28s} rel {T._rel(o, ref_nchw):.4f}  |ref| {ref_nchw.norm().item():.3e}")
cmp("z_conv9_2 (g_c9)", B["g_c9"], Z["z_conv9_2"].grad)
for i in range(9, 5, -1):
    cmp(f"z_conv{i}_1 (gx_conv{i}_2_0)", B[f"gx_conv{i}_2_0"], Z[f"z_conv{i}_1"].grad)
    cmp(f"u_upv{i} (gx_conv{i}_1_0)", B[f"gx_conv{i}_1_0"], Z[f"u_upv{i}"].grad)
    prev = f"conv{i-1}_2" if i > 6 else "conv5_2"
    cmp(f"z_{prev} (gx_upv{i})", B[f"gx_upv{i}"], Z[f"z_{prev}"].grad)
for i in range(5, 0, -1):
    cmp(f"z_conv{i}_1 (gx_conv{i}_2_0)", B[f"gx_conv{i}_2_0"], Z[f"z_conv{i}_1"].grad)
    if i > 1:
        cmp(f"z_conv{i-1}_2 (g_c{i-1})", B[f"g_c{i-1}"], Z[f"z_conv{i-1}_2"].grad)
