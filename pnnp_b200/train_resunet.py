"""Synthetic-pair training step (T1) for ResUnet on the B200 kernels — the same loop body as train.UNetTrainStep
(trainer_SID.py:93-101: pred = net(lr); F.l1_loss(pred.clamp(0,1), hr); backward; Adam) for archs/ResUnet.py:46-88.

Explicit backward on the kernels the UNet step uses:
  * ResidualBlock (modules.py:176-197: conv(no bias)+ReLU -> conv(no bias), += shortcut):  weight gradients = pnnp_wgrad_nhwc;
    data gradients = the tcgen05 conv kernel on transposed + flipped weights with ReLU' fused as the epilogue mask and the
    shortcut's gradient fused as the epilogue residual (identity shortcut: the incoming gradient itself; 1x1 shortcut: its own
    1x1 data-gradient conv);
  * stride-2 down-sampling conv (modules.py:130-138): the output gradient is zero-inserted onto the input grid, after which its
    data AND weight gradients are those of a stride-1 3x3 conv (zeros contribute nothing) — 4x the minimal tensor work, no new
    kernel;
  * ConvTranspose2d, the 1x1 head, bias sums, L1 loss, Adam: as in the UNet step.
The 1x1 shortcut's weight gradient is the centre tap of a 3x3 weight-gradient launch.  Weight re-packing after Adam goes through
the network's pack cache (framework ops), the step runs eagerly: functional parity first, this step is not tuned like the UNet's
(no batched packing, no graph replay).  Under DDP the gradient scratch is all-reduced in one piece after the backward pass.
"""
import torch

from . import _lib, distributed as D
from .archs import ResUnet, _conv, _pad16, _to_nhwc16
from .train import _Scratch, _desc

L = _lib


def _pack(w3, dtype=torch.bfloat16):
    """[taps, rows, cin] fp32 -> zero-padded [taps][rows_pad16][cin_pad16] bf16 (the conv kernel's weight layout)."""
    t, r, c = w3.shape
    buf = torch.zeros((t, _pad16(r), _pad16(c)), dtype=dtype, device=w3.device)
    buf[:, :r, :c] = w3.to(dtype)
    return buf


class ResUnetTrainStep:
    """forward + L1 loss + backward + Adam for a ResUnet on one GPU (one rank of a DDP job)."""

    def __init__(self, net: ResUnet, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.net = net
        self.device = next(net.parameters()).device
        _lib.require_cuda_device(self.device, "the training step")
        self.betas, self.eps, self.t = betas, eps, 0
        self.adam_state = torch.tensor([float(lr), 0.0], dtype=torch.float32, device=self.device)
        self._lr = float(lr)
        self.use_graph = False
        params = list(net.named_parameters())
        total = sum(p.numel() for _, p in params)
        self.flat_p = torch.empty(total, dtype=torch.float32, device=self.device)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.m, self.v = torch.zeros_like(self.flat_p), torch.zeros_like(self.flat_p)
        self.slices, off = {}, 0
        for name, p in params:
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[off:off + n].view_as(p)
            self.slices[name] = (off, n, tuple(p.shape))
            off += n
        if D.world()[1] > 1:
            torch.distributed.broadcast(self.flat_p, 0)              # every rank starts from rank 0's weights
        self.scr = _Scratch(self.device)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        # gradient scratch: [taps][ci_pad16][co] per conv weight (1x1 shortcuts get a 3x3 region, centre tap used), bias sums behind
        self.dw, self.db, total_dw, unpack = {}, {}, 0, []
        for name, m in net.named_modules():
            if isinstance(m, torch.nn.ConvTranspose2d):
                shape = (4, m.weight.shape[0], m.weight.shape[1])
            elif isinstance(m, torch.nn.Conv2d):
                shape = (1, m.weight.shape[0], m.weight.shape[1]) if name == "conv10" else (9, _pad16(m.weight.shape[1]), m.weight.shape[0])
            else:
                continue
            self.dw[name] = (total_dw, shape)
            total_dw += shape[0] * shape[1] * shape[2]
            if m.bias is not None:
                self.db[name] = (total_dw, m.bias.numel())
                total_dw += (m.bias.numel() + 3) // 4 * 4
        self.dw_flat = torch.zeros(total_dw + 64, dtype=torch.float32, device=self.device)      # + a dummy bias-sum target
        self._dummy = self.dw_flat[total_dw:total_dw + 64]
        gp = lambda pname: self.flat_g.data_ptr() + 4 * self.slices[pname][0]
        for name, m in net.named_modules():
            if name not in self.dw:
                continue
            src = self._dw_view(name)
            if isinstance(m, torch.nn.ConvTranspose2d):              # grad[ci][co][tap] = dw[tap][ci][co]
                ci, co = m.weight.shape[0], m.weight.shape[1]
                unpack.append(_desc(src.data_ptr(), gp(name + ".weight"), 0, (ci, co, 4), (co, 1, ci * co), (co * 4, 4, 1)))
            elif name == "conv10":
                unpack.append(_desc(src.data_ptr(), gp(name + ".weight"), 0, (m.weight.numel(),), (1,), (1,)))
            else:
                co, cin, k = m.weight.shape[0], m.weight.shape[1], m.weight.shape[2]
                ci_total = self.dw[name][1][1]
                if k == 3:                                           # grad[co][ci][tap] = dw[tap][ci][co]
                    unpack.append(_desc(src.data_ptr(), gp(name + ".weight"), 0, (co, cin, 9), (1, co, ci_total * co), (cin * 9, 9, 1)))
                else:                                                # 1x1 shortcut: centre tap of the 3x3 region
                    unpack.append(_desc(src.data_ptr() + 4 * 4 * ci_total * co, gp(name + ".weight"), 0, (co, cin), (1, co), (cin, 1)))
            if name in self.db:
                o, nb = self.db[name]
                unpack.append(_desc(self.dw_flat.data_ptr() + 4 * o, gp(name + ".bias"), 0, (nb,), (1,), (1,)))
        arr = (_lib.CopyDesc * len(unpack))(*unpack)
        self._unpack_tab = (torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device), len(unpack))

    # ---------------------------------------------------------------- helpers
    def close(self):
        if torch.device(self.device).type == "cuda":
            torch.cuda.synchronize(self.device)

    def _stream(self):
        return _lib.stream_ptr(self.device)

    def _dw_view(self, name):
        off, shape = self.dw[name]
        return self.dw_flat[off:off + shape[0] * shape[1] * shape[2]].view(shape)

    def _db_view(self, name):
        off, n = self.db[name]
        return self.dw_flat[off:off + n]

    def _grad_view(self, name):
        off, n, shape = self.slices[name]
        return self.flat_g[off:off + n].view(shape)

    def _w(self, name):
        return self.net.get_submodule(name).weight.detach()

    def _bias_sum(self, g, name):
        pixels = g.numel() // g.shape[-1]
        L.check(L.lib().pnnp_act_bwd_bias(g.data_ptr(), None, self._db_view(name).data_ptr(), pixels, g.shape[-1], _lib.ACT_NONE,
                                          self._stream()), "bias gradient")

    def _wgrad(self, mode, g, x, name, ci_off=0):
        co = g.shape[-1]
        n, h, w, ci = x.shape
        dw = self._dw_view(name)
        L.check(L.lib().pnnp_wgrad_nhwc(mode, g.data_ptr(), co, co, x.data_ptr(), ci, ci, n, h, w, dw.data_ptr(), ci_off, dw.shape[1],
                                        dw.shape[2], self._stream()), "wgrad_nhwc")

    def _dgrad3(self, g, w4, out_name, c_lo, c_hi, mask=None, resid=None):
        """Data gradient of a 3x3 s1 conv w.r.t. input channels [c_lo, c_hi): conv of g with W^T flipped by 180 degrees."""
        n, h, w, co = g.shape
        wd = _pack(w4[:, c_lo:c_hi].flip(2, 3).permute(2, 3, 1, 0).reshape(9, c_hi - c_lo, co))      # [8 - tap][ci][co]
        gx = self.scr.get(out_name, (n, h, w, c_hi - c_lo))
        _conv(_lib.CONV3, g, wd, None, gx, c_hi - c_lo, _lib.ACT_NONE, mask=mask, mask_slope=0.0, resid=resid)
        return gx

    # ---------------------------------------------------------------- forward (every intermediate kept)
    def forward(self, x):
        net, nf = self.net, self.net.nf
        x = x.float().contiguous()
        n, c, h, w = x.shape
        if h % 16 or w % 16:
            raise RuntimeError("pnnp_b200: h and w must be multiples of 16")
        s = {}
        buf = lambda name, hh, ww, cc: self.scr.get("a_" + name, (n, hh, ww, cc))
        RELU, NONE = _lib.ACT_RELU, _lib.ACT_NONE
        pk = lambda name, kind="conv": net._packed(name, kind)
        s["x16"] = _to_nhwc16(x, buf("x16", h, w, 16))
        wi, bi = pk("conv_in")
        cur = s["cin"] = _conv(_lib.CONV3, s["x16"], wi, bi, buf("cin", h, w, nf), nf, RELU)
        hh, ww = h, w
        for i in range(1, 6):
            co = nf * 2 ** (i - 1)
            s[f"in{i}"] = cur
            s[f"t{i}"] = _conv(_lib.CONV3, cur, pk(f"conv{i}.block.0.conv.conv")[0], None, buf(f"t{i}", hh, ww, co), co, RELU)
            s[f"b{i}"] = cur = _conv(_lib.CONV3, s[f"t{i}"], pk(f"conv{i}.block.1.conv.conv")[0], None, buf(f"b{i}", hh, ww, co), co, NONE,
                                     resid=cur)
            if i < 5:
                wp, bp = pk(f"pool{i}.conv")
                cur = _conv(_lib.CONV3S2, cur, wp, bp, buf(f"d{i}", hh // 2, ww // 2, co * 2), co * 2, NONE)
                hh, ww = hh // 2, ww // 2
        for i in range(6, 10):
            co = nf * 2 ** (9 - i)
            s[f"upin{i}"] = cur
            wu, bu = pk(f"upv{i}", "convT")
            up = s[f"u{i}"] = _conv(_lib.CONVT, cur, wu, bu, buf(f"u{i}", hh * 2, ww * 2, co), co, NONE)
            hh, ww = hh * 2, ww * 2
            skip = s[f"b{10 - i}"]
            s[f"t{i}"] = _conv(_lib.CONV3, up, pk(f"conv{i}.block.0.conv.conv")[0], None, buf(f"t{i}", hh, ww, co), co, RELU, x1=skip)
            sc = _conv(_lib.CONV1, up, pk(f"conv{i}.short_cut.0.conv.conv")[0], None, buf(f"s{i}", hh, ww, co), co, NONE, x1=skip)
            s[f"b{i}"] = cur = _conv(_lib.CONV3, s[f"t{i}"], pk(f"conv{i}.block.1.conv.conv")[0], None, buf(f"b{i}", hh, ww, co), co, NONE,
                                     resid=sc)
        w10, b10 = pk("conv10")
        pred = self.scr.get("pred", (n, net.out_nc, h, w), torch.float32)
        _conv(_lib.CONV1, cur, w10, b10, pred, net.out_nc, NONE, out_mode=_lib.OUT_NCHW_F32, resid_nchw=x if net.res else None)
        return pred, s

    # ---------------------------------------------------------------- backward
    def _block_bwd(self, i, g_b, srcs):
        """ResidualBlock i: g_b = gradient w.r.t. the block's output; srcs = its input(s).  Returns the input gradients."""
        name, co = f"conv{i}", g_b.shape[-1]
        t = self.scr.bufs[f"a_t{i}"]
        self._wgrad(0, g_b, t, f"{name}.block.1.conv.conv")
        g_t = self._dgrad3(g_b, self._w(f"{name}.block.1.conv.conv"), f"g_t{i}", 0, co, mask=t)          # ReLU'(t) fused
        w1 = self._w(f"{name}.block.0.conv.conv")
        outs, c_off = [], 0
        for k, src in enumerate(srcs):
            ck = src.shape[-1]
            self._wgrad(0, g_t, src, f"{name}.block.0.conv.conv", ci_off=c_off)
            if len(srcs) == 1:
                sc_grad = g_b                                                                            # identity shortcut
            else:                                                                                        # 1x1 shortcut conv
                self._wgrad(0, g_b, src, f"{name}.short_cut.0.conv.conv", ci_off=c_off)                   # centre tap = the 1x1 gradient
                wsc = self._w(f"{name}.short_cut.0.conv.conv")[:, c_off:c_off + ck, 0, 0]                # [co][ck]
                sc_grad = self.scr.get(f"g_sc{i}_{k}", tuple(src.shape))
                _conv(_lib.CONV1, g_b, _pack(wsc.t().reshape(1, ck, co)), None, sc_grad, ck, _lib.ACT_NONE)
            mask = self.scr.bufs["a_cin"] if (i == 1 and k == 0) else None                               # conv_in's ReLU' on its output
            outs.append(self._dgrad3(g_t, w1, f"g_in{i}_{k}", c_off, c_off + ck, mask=mask, resid=sc_grad))
            c_off += ck
        return outs

    def backward(self, gpred, s):
        net, nf = self.net, self.net.nf
        n, _, h, w = gpred.shape
        self.flat_g.zero_()
        self.dw_flat.zero_()
        g = self.scr.get("g_b9", (n, h, w, nf))
        w10 = self._w("conv10").reshape(net.out_nc, nf).to(torch.bfloat16).float()
        L.check(L.lib().pnnp_head_bwd(gpred.data_ptr(), s["b9"].data_ptr(), w10.data_ptr(), g.data_ptr(), self._dw_view("conv10").data_ptr(),
                                      self._db_view("conv10").data_ptr(), self._dummy.data_ptr(), n, h, w, nf, net.out_nc, _lib.ACT_NONE,
                                      self._stream()), "head_bwd")
        g_skip = {}
        for i in range(9, 5, -1):
            g_up, g_skip[10 - i] = self._block_bwd(i, g, [s[f"u{i}"], s[f"b{10 - i}"]])
            m = net.get_submodule(f"upv{i}")
            ci, co = m.weight.shape[0], m.weight.shape[1]
            self._bias_sum(g_up, f"upv{i}")
            self._wgrad(1, g_up, s[f"upin{i}"], f"upv{i}")
            x_in = s[f"upin{i}"]
            g = self.scr.get(f"g_upin{i}", tuple(x_in.shape))
            wd = _pack(m.weight.detach().permute(2, 3, 0, 1).reshape(4, ci, co))                          # [a*2+b][ci][co]
            _conv(_lib.CONV2S2, g_up, wd, None, g, ci, _lib.ACT_NONE)
        for i in range(5, 0, -1):
            if i < 5:                                                # stride-2 conv pool{i}: zero-insert its output gradient
                b_i = s[f"b{i}"]
                nb, hb, wb, cb = b_i.shape
                G = self.scr.get(f"G{i}", (nb, hb, wb, g.shape[-1]))
                G.zero_()
                G[:, 0::2, 0::2, :] = g
                self._bias_sum(g, f"pool{i}.conv")
                self._wgrad(0, G, b_i, f"pool{i}.conv")
                g = self._dgrad3(G, self._w(f"pool{i}.conv"), f"g_b{i}", 0, cb, resid=g_skip[i])          # + the skip connection's gradient
            (g,) = self._block_bwd(i, g, [s[f"in{i}"]])
        self._bias_sum(g, "conv_in")                                 # g already carries conv_in's ReLU'
        self._wgrad(0, g, s["x16"], "conv_in")
        if D.world()[1] > 1:
            torch.distributed.all_reduce(self.dw_flat, op=torch.distributed.ReduceOp.SUM)
        tab, nd = self._unpack_tab
        L.check(L.lib().pnnp_strided_copy_batch(tab.data_ptr(), nd, 48, self._stream()), "unpack gradients")

    # ---------------------------------------------------------------- the step
    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        if float(value) != self._lr:
            self._lr = float(value)
            self.adam_state[0:1].fill_(self._lr)

    def sync_parameters(self, src=0):
        if D.world()[1] > 1:
            torch.distributed.broadcast(self.flat_p, src)
        self.refresh_packed()

    def refresh_packed(self):
        self.net.__dict__.get("_pack_cache", {}).clear()             # the kernels wrote the weights in place: re-pack on next use

    def step(self, lr_crops, hr_crops, grad_allreduce=True):
        """One optimisation step on (noisy, clean) crops (CUDA fp32 NCHW).  Returns the loss as a 0-d CUDA tensor."""
        self.t += 1
        pred, saved = self.forward(lr_crops)
        hr = hr_crops.float().contiguous()
        gpred = self.scr.get("gpred", tuple(pred.shape), torch.float32)
        L.check(L.lib().pnnp_l1_loss(pred.data_ptr(), hr.data_ptr(), gpred.data_ptr(), pred.numel(), self.loss_sum.data_ptr(),
                                     self._stream()), "l1_loss")
        self.backward(gpred, saved)
        gscale = 1.0 / D.world()[1] if D.world()[1] > 1 else 1.0
        L.check(L.lib().pnnp_adam_step_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                           self.flat_p.numel(), self.adam_state.data_ptr(), self.betas[0], self.betas[1], self.eps,
                                           gscale, self._stream()), "adam_step")
        self.refresh_packed()
        return (self.loss_sum / pred.numel()).float()[0]
