"""Synthetic-pair training step (T1) for UNetSeeInDark on the B200 kernels.

Reference loop body (trainer_SID.py:93-101):  pred = net(lr); loss = F.l1_loss(pred.clamp(0,1), hr)
(losses/base_loss.py:92-103); loss.backward(); Adam(lr=1e-4).step()  (trainer_SID.py:44).

Here the backward pass is explicit (no autograd graph, no eager fallback):
  * data gradients of the 3x3 convs  = the SAME tcgen05 conv kernel with transposed + flipped weights;
    ConvTranspose2d data gradient    = its 2x2 stride-2 mode;
  * weight gradients                 = the tcgen05 split-K GEMM of csrc/wgrad_nhwc_tc.cu straight from the NHWC tensors
    (MN-major operands via TMA: K = pixels, zero padding = out-of-bounds fill, no transposed copies);
  * LeakyReLU', bias gradients, max-pool routing (+ skip-connection add), the 1x1 head, the L1 loss and Adam
    are the CUDA-core kernels of csrc/train_kernels.cu.
Parameters, gradients and Adam moments live in flat fp32 buffers (one all-reduce per step under DDP);
activations and activation gradients are NHWC bf16.
"""
import os

import torch

from . import _lib, distributed as D
from .archs import UNetSeeInDark, _conv, _pad16, _to_nhwc16

L = _lib


def _pack_conv_weight(w4):
    """[rows, cin, k, k] fp32 -> kernel layout [k*k][rows_pad16][cin_pad16] bf16."""
    rows, cin, k = w4.shape[0], w4.shape[1], w4.shape[2]
    buf = torch.zeros((k * k, _pad16(rows), _pad16(cin)), dtype=torch.bfloat16, device=w4.device)
    buf[:, :rows, :cin] = w4.permute(2, 3, 0, 1).reshape(k * k, rows, cin).to(torch.bfloat16)
    return buf


class _StaticPacked:
    """Entry of the network's pack cache whose bf16 buffer is refreshed by the step's batched pack launch instead of being
    rebuilt with framework ops; a framework-side in-place change of the weight (load_state_dict) triggers a refresh."""

    def __init__(self, owner, module, weight_buf):
        self.owner, self.module, self.weight = owner, module, weight_buf
        self.version = module.weight._version

    def get(self, device):
        if self.module.weight._version != self.version:
            self.owner.refresh_packed()
        b = self.module.bias
        return self.weight, (None if b is None else b.detach())


def archs_x_mode():
    from . import archs
    return archs._X_MODE


def _desc(src_ptr, dst_ptr, dst_bf16, dims, sstrides, dstrides):
    d = _lib.CopyDesc()
    d.src, d.dst, d.dst_bf16 = src_ptr, dst_ptr, int(dst_bf16)
    dims, sstrides, dstrides = list(dims), list(sstrides), list(dstrides)
    while len(dims) < 4:                               # leading unit dimensions
        dims.insert(0, 1); sstrides.insert(0, 0); dstrides.insert(0, 0)
    for i in range(4):
        d.dim[i], d.sstride[i], d.dstride[i] = int(dims[i]), int(sstrides[i]), int(dstrides[i])
    return d


def _split_desc(d, max_elems=131072):
    """One launch of the batched copy gives every descriptor the same number of blocks, so a 2.4 M-element weight tensor next to a
    2 K-element one left most of the GPU idle (r02 launch list: 0.18 ms per step for 31 MB).  Descriptors are cut along their
    leading dimensions into pieces of at most `max_elems` elements."""
    dims, ss, ds = list(d.dim), list(d.sstride), list(d.dstride)
    esz_d = 2 if d.dst_bf16 else 4
    pieces = [(int(d.src), int(d.dst), dims)]
    for ax in range(3):
        out = []
        for src, dst, dm in pieces:
            n = dm[0] * dm[1] * dm[2] * dm[3]
            if n <= max_elems or dm[ax] == 1:
                out.append((src, dst, dm))
                continue
            chunk = max(1, dm[ax] // -(-n // max_elems))           # as few chunks along this axis as bring a piece under the limit
            for i in range(0, dm[ax], chunk):
                nd = list(dm)
                nd[ax] = min(chunk, dm[ax] - i)
                out.append((src + 4 * i * ss[ax], dst + esz_d * i * ds[ax], nd))
        pieces = out
    res = []
    for src, dst, dm in pieces:
        p = _lib.CopyDesc()
        p.src, p.dst, p.dst_bf16 = src, dst, d.dst_bf16
        for i in range(4):
            p.dim[i], p.sstride[i], p.dstride[i] = dm[i], ss[i], ds[i]
        res.append(p)
    return res


class _Scratch:
    """Named device buffers reused across steps."""

    def __init__(self, device):
        self.device, self.bufs = device, {}

    def get(self, name, shape, dtype=torch.bfloat16):
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t


class UNetTrainStep:
    """forward + L1 loss + backward + Adam for a UNetSeeInDark on one GPU (one rank of a DDP job)."""

    def __init__(self, net: UNetSeeInDark, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.net = net
        self.device = next(net.parameters()).device
        _lib.require_cuda_device(self.device, "the training step")
        self.betas, self.eps, self.t = betas, eps, 0
        # learning rate and step count in device memory: the whole step is replayed as one CUDA graph (see step())
        self.adam_state = torch.tensor([float(lr), 0.0], dtype=torch.float32, device=self.device)
        self._lr = float(lr)
        # CUDA-graph replay of the step, single process and DDP alike (the bucketed all-reduces are captured with it).  Before the
        # process group is destroyed the graphs have to go (close()): r01's destroy_process_group() never returned with captured
        # all-reduces alive; with the graphs dropped first it returns in 0.5 s (r02, tools/ddp_check.py).  PNNP_TRAIN_GRAPH=0: eager.
        self.use_graph = os.environ.get("PNNP_TRAIN_GRAPH", "1") != "0"
        self._graphs = {}
        # flat fp32 parameter / gradient / moment buffers; the module's parameters become views of the flat buffer
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
        self.sync_parameters()
        self.scr = _Scratch(self.device)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        # Gradient scratch of every layer, carved out of ONE buffer (a single memset per step) IN BACKWARD ORDER, so that what the
        # backward pass finishes first lies first: [taps][ci_pad16][co] fp32 weight gradient (the layout the wgrad kernel adds into
        # with coalesced reds) followed by the layer's [co] bias gradient; the 1x1 head (weight [out_nc][nf] + bias) leads.  Under
        # DDP the buffer is all-reduced in contiguous BUCKETS while the rest of the backward pass still runs (SURVEY 8e:
        # "bucketed and overlapped with backward"); `unpack` then scatters it into the parameters' own layout in flat_g.
        mods = dict(net.named_modules())
        order = ["conv10_1"]
        for i in range(9, 5, -1):
            order += [f"conv{i}_2", f"conv{i}_1", f"upv{i}"]
        for i in range(5, 0, -1):
            order += [f"conv{i}_2", f"conv{i}_1"]
        self.dw, self.db, total_dw = {}, {}, 0
        for name in order:
            m = mods[name]
            if isinstance(m, torch.nn.ConvTranspose2d):
                shape = (4, m.weight.shape[0], m.weight.shape[1])
            elif m.kernel_size == (3, 3):
                shape = (9, _pad16(m.weight.shape[1]), m.weight.shape[0])
            else:
                shape = (1, m.weight.shape[0], m.weight.shape[1])                 # conv10_1: [1][out_nc][nf], the parameter's own layout
            self.dw[name] = (total_dw, shape)
            total_dw += shape[0] * shape[1] * shape[2]
            if m.bias is not None:
                self.db[name] = (total_dw, m.bias.numel())
                total_dw += (m.bias.numel() + 3) // 4 * 4                         # keep every region 16-byte aligned
        self.dw_flat = torch.zeros(total_dw, dtype=torch.float32, device=self.device)
        # bucket = contiguous slice of dw_flat closed by the layer whose gradients complete it (~5-9 MB each; the last, tiny one is
        # the only part of the reduction that cannot overlap with anything)
        closers = ("conv6_2", "upv6", "conv5_2", "conv4_1", "conv1_1")
        self.buckets, start = {}, 0
        for name in order:
            if name in closers:
                end = (self.db[name][0] + (self.db[name][1] + 3) // 4 * 4) if name in self.db else self.dw[name][0] + \
                    self.dw[name][1][0] * self.dw[name][1][1] * self.dw[name][1][2]
                self.buckets[name] = (start, end)
                start = end
        assert start == total_dw
        self.comm_stream = torch.cuda.Stream(self.device) if D.world()[1] > 1 else None
        self._build_copy_tables()

    def close(self):
        """Drop the captured graphs (they hold NCCL work under DDP) — call before torch.distributed.destroy_process_group()."""
        self._graphs.clear()
        if torch.device(self.device).type == "cuda":
            torch.cuda.synchronize(self.device)

    def sync_parameters(self, src=0):
        """Every rank starts from (and, after a checkpoint reload, continues from) rank `src`'s parameters, as the
        DistributedDataParallel constructor does: each process seeds its own CPU generator, so `initialize_weights` and the
        default-initialised ConvTranspose2d biases differ between ranks until this broadcast.  No-op for a single process."""
        if D.world()[1] > 1:
            torch.distributed.broadcast(self.flat_p, src)
            if hasattr(self, "_pack_tab"):
                self.refresh_packed()

    # ---------------------------------------------------------------- batched weight packing / gradient layout
    def _build_copy_tables(self):
        """Descriptor tables of the two batched strided-copy launches of a step (csrc/train_kernels.cu):
        `pack`   fp32 master weights -> every bf16 tensor-core layout the step reads (forward layers as the network's pack
                 cache holds them, transposed + flipped data-gradient forms), run after each Adam update;
        `unpack` wgrad scratch [tap][ci][co] -> the parameters' own gradient layout in the flat gradient buffer."""
        net = self.net
        with torch.no_grad():                              # dry run: lets the network decide each layer's kernel mode / layout
            self.forward(torch.zeros((1, net.conv1_1.weight.shape[1], 32, 32), device=self.device))
        self.scr.bufs.clear()
        fp = lambda name: self.flat_p.data_ptr() + 4 * self.slices[name + ".weight"][0]
        gp = lambda name: self.flat_g.data_ptr() + 4 * self.slices[name + ".weight"][0]
        pack, unpack, self.wd = [], [], {}
        cache = net.__dict__["_pack_cache"]
        for (name, kind), entry in list(cache.items()):
            m, buf = net.get_submodule(name), entry.weight
            rows_p, cin_p = buf.shape[1], buf.shape[2]
            if kind == "convT":                            # W[ci][co][a][b] -> [a*2+b][co][ci]
                ci, co = m.weight.shape[0], m.weight.shape[1]
                pack.append(_desc(fp(name), buf.data_ptr(), 1, (4, co, ci), (1, 4, co * 4), (rows_p * cin_p, cin_p, 1)))
            elif kind == "conv3x":                         # W[co][ci][ky][kx] -> [ky][kx*co + co'][ci]
                co, ci = m.weight.shape[0], m.weight.shape[1]
                pack.append(_desc(fp(name), buf.data_ptr(), 1, (3, 3, co, ci), (3, 1, ci * 9, 9), (rows_p * cin_p, co * cin_p, cin_p, 1)))
            else:                                          # W[co][ci][ky][kx] -> [tap][co][ci]
                co, ci, kk = m.weight.shape[0], m.weight.shape[1], m.weight.shape[2] * m.weight.shape[3]
                pack.append(_desc(fp(name), buf.data_ptr(), 1, (kk, co, ci), (1, ci * kk, kk), (rows_p * cin_p, cin_p, 1)))
            cache[(name, kind)] = _StaticPacked(self, m, buf)
        for name, m in net.named_modules():
            if isinstance(m, torch.nn.ConvTranspose2d):    # dgrad: [a*2+b][ci][co] = W[ci][co][a][b];  grad[ci][co][tap] = dw[tap][ci][co]
                ci, co = m.weight.shape[0], m.weight.shape[1]
                buf = torch.zeros((4, _pad16(ci), _pad16(co)), dtype=torch.bfloat16, device=self.device)
                self.wd[name] = buf
                pack.append(_desc(fp(name), buf.data_ptr(), 1, (4, ci, co), (1, co * 4, 4), (buf.shape[1] * buf.shape[2], buf.shape[2], 1)))
                unpack.append(_desc(self._dw_view(name).data_ptr(), gp(name), 0, (ci, co, 4), (co, 1, ci * co), (co * 4, 4, 1)))
            elif isinstance(m, torch.nn.Conv2d) and m.kernel_size == (3, 3):
                co, cin = m.weight.shape[0], m.weight.shape[1]
                ci_total = self.dw[name][1][1]
                unpack.append(_desc(self._dw_view(name).data_ptr(), gp(name), 0, (co, cin, 9), (1, co, ci_total * co), (cin * 9, 9, 1)))
                if name == "conv1_1":
                    continue                               # the network input needs no gradient
                halves = (cin // 2, cin // 2) if (name.endswith("_1") and int(name[4]) >= 6) else (cin,)   # cat([up, skip], 1)
                c_off = 0
                for k, ck in enumerate(halves):            # dgrad: [8 - tap][ci][co] <- W[co][c_off + ci][tap]  (180-degree flip)
                    if ck <= 32 and archs_x_mode():        # narrow output: the x-shift-in-N layout [ky][kx*ck + ci][co], as in the forward
                        buf = torch.zeros((3, _pad16(3 * ck), _pad16(co)), dtype=torch.bfloat16, device=self.device)
                        pack.append(_desc(fp(name) + 4 * (c_off * 9 + 8), buf.data_ptr(), 1, (3, 3, ck, co), (-3, -1, 9, cin * 9),
                                          (buf.shape[1] * buf.shape[2], ck * buf.shape[2], buf.shape[2], 1)))
                    else:
                        buf = torch.zeros((9, _pad16(ck), _pad16(co)), dtype=torch.bfloat16, device=self.device)
                        pack.append(_desc(fp(name) + 4 * (c_off * 9 + 8), buf.data_ptr(), 1, (9, ck, co), (-1, 9, cin * 9),
                                          (buf.shape[1] * buf.shape[2], buf.shape[2], 1)))
                    self.wd[(name, k)] = buf
                    c_off += ck
        for layer, (off, nb) in self.db.items():            # bias gradients: reduce buffer -> flat gradient buffer
            src = self.dw_flat.data_ptr() + 4 * off
            unpack.append(_desc(src, self.flat_g.data_ptr() + 4 * self.slices[layer + ".bias"][0], 0, (nb,), (1,), (1,)))
        n10 = self.slices["conv10_1.weight"][1]              # 1x1 head weight gradient: same layout on both sides
        unpack.append(_desc(self._dw_view("conv10_1").data_ptr(), gp("conv10_1"), 0, (n10,), (1,), (1,)))

        def table(descs):
            descs = [piece for d in descs for piece in _split_desc(d)]
            arr = (_lib.CopyDesc * len(descs))(*descs)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            return host.to(self.device), len(descs)
        self._pack_tab, self._unpack_tab = table(pack), table(unpack)
        self.refresh_packed()

    def refresh_packed(self):
        """Re-pack every bf16 weight layout from the fp32 master weights (one launch)."""
        tab, n = self._pack_tab
        L.check(L.lib().pnnp_strided_copy_batch(tab.data_ptr(), n, 24, self._stream()), "pack weights")
        cache = self.net.__dict__["_pack_cache"]
        for key in list(cache):
            if isinstance(cache[key], _StaticPacked):
                cache[key].version = cache[key].module.weight._version
            else:
                del cache[key]                             # framework-packed entries cannot see this step's in-place update

    # ---------------------------------------------------------------- small wrappers over the C ABI
    def _grad_view(self, name):
        """Where the backward kernels write the gradient of parameter `name`: bias gradients and the 1x1 head's weight gradient
        go straight into the reduce buffer (same layout as the parameter); 3x3 / transposed-conv weights use _dw_view."""
        layer, kind = name.rsplit(".", 1)
        if kind == "bias" and layer in self.db:
            off, n = self.db[layer]
            return self.dw_flat[off:off + n]
        if name == "conv10_1.weight":
            return self._dw_view("conv10_1").view(self.slices[name][2])
        off, n, shape = self.slices[name]
        return self.flat_g[off:off + n].view(shape)

    def _bucket_done(self, layer, grad_allreduce):
        """DDP: `layer` completes a bucket of the reduce buffer -> all-reduce(SUM) it on the communication stream while the current
        stream goes on with the backward pass (the 1 / world factor is folded into Adam)."""
        if not grad_allreduce or self.comm_stream is None or layer not in self.buckets:
            return
        a, b = self.buckets[layer]
        cur = torch.cuda.current_stream(self.device)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            torch.distributed.all_reduce(self.dw_flat[a:b], op=torch.distributed.ReduceOp.SUM)

    def _dw_view(self, name):
        off, shape = self.dw[name]
        return self.dw_flat[off:off + shape[0] * shape[1] * shape[2]].view(shape)

    def _stream(self):
        return _lib.stream_ptr(self.device)

    def _act_bwd(self, g, out, dbias, act):
        pixels = g.numel() // g.shape[-1]
        L.check(L.lib().pnnp_act_bwd_bias(g.data_ptr(), _lib.ptr(out), _lib.ptr(dbias), pixels, g.shape[-1], act, self._stream()),
                "act_bwd_bias")

    def _wgrad_nhwc(self, mode, g, co, x, dw, ci_off, ci_total):
        """dw[tap][ci_off + ci][co] += the weight gradient of a 3x3 conv (mode 0) / 2x2 stride-2 transposed conv (mode 1)
        with pre-activation output gradient g and input x (both NHWC bf16), by the tcgen05 kernel of csrc/wgrad_nhwc_tc.cu."""
        n, h, w, ci = x.shape
        L.check(L.lib().pnnp_wgrad_nhwc(mode, g.data_ptr(), co, g.shape[-1], x.data_ptr(), ci, ci, n, h, w, dw.data_ptr(), ci_off,
                                        ci_total, dw.shape[-1], self._stream()), "wgrad_nhwc")

    # ---------------------------------------------------------------- layer backward passes
    def _conv3_bwd(self, name, g, srcs, need_dx, dx_masks=None, bias_done=False):
        """g: NHWC bf16 gradient w.r.t. the PRE-activation output of conv `name` (the producer of g fused this layer's
        LeakyReLU').  srcs: list of NHWC bf16 inputs (1, or 2 for torch.cat([up, skip], 1)).  dx_masks[k]: activated tensor
        whose LeakyReLU' is fused into the k-th input gradient (None: the input is not an activation output).
        Returns the input gradients."""
        m = self.net.get_submodule(name)
        co = m.weight.shape[0]
        n, h, w, _ = g.shape
        if not bias_done:
            self._act_bwd(g, None, self._grad_view(name + ".bias"), _lib.ACT_NONE)    # bias gradient = sum over pixels
        ci_total = sum(s.shape[-1] for s in srcs)
        dw = self._dw_view(name)
        c_off = 0
        for s in srcs:
            self._wgrad_nhwc(0, g, co, s, dw, c_off, ci_total)
            c_off += s.shape[-1]
        self._bucket_done(name, getattr(self, "_ar", False))  # this layer's gradients are complete: the data gradient overlaps the reduce
        gxs = []
        if need_dx:
            for k, s in enumerate(srcs):                  # data gradient: the forward kernel on the transposed + flipped weights
                ck = s.shape[-1]
                gx = self.scr.get(f"gx_{name}_{k}", (n, h, w, ck))
                wd = self.wd[(name, k)]
                _conv(_lib.CONV3X if wd.shape[0] == 3 else _lib.CONV3, g, wd, None, gx, ck, _lib.ACT_NONE,
                      mask=None if dx_masks is None else dx_masks[k])
                gxs.append(gx)
        return gxs

    def _convT_bwd(self, name, g_up, x_in):
        """ConvTranspose2d(2, stride 2) backward: g_up NHWC bf16 [n,2h,2w,co] -> g_in [n,h,w,ci] (times LeakyReLU'(x_in): x_in is an
        activation output); dW [ci][co][2][2]; db."""
        m = self.net.get_submodule(name)
        ci, co = m.weight.shape[0], m.weight.shape[1]
        n, h, w, _ = x_in.shape
        self._act_bwd(g_up, None, self._grad_view(name + ".bias"), _lib.ACT_NONE)
        dw = self._dw_view(name)
        self._wgrad_nhwc(1, g_up, co, x_in, dw, 0, ci)
        self._bucket_done(name, getattr(self, "_ar", False))
        gx = self.scr.get("gx_" + name, (n, h, w, ci))
        _conv(_lib.CONV2S2, g_up, self.wd[name], None, gx, ci, _lib.ACT_NONE, mask=x_in)      # [a*2+b][ci][co] = W[ci][co][a][b]
        return gx

    # ---------------------------------------------------------------- the step
    def forward(self, x):
        """Training forward: like UNetSeeInDark.forward but every intermediate is kept (no fused head)."""
        net, nf = self.net, self.net.nf
        x = x.float().contiguous()
        n, c, h, w = x.shape
        if h % 16 or w % 16:
            raise RuntimeError("pnnp_b200: h and w must be multiples of 16")
        s = {}
        buf = lambda name, hh, ww, cc: self.scr.get("a_" + name, (n, hh, ww, cc))
        LK = _lib.ACT_LEAKY
        cur = _to_nhwc16(x, buf("x16", h, w, 16))
        s["x16"] = cur
        hh, ww = h, w
        for i in range(1, 6):
            co = nf * 2 ** (i - 1)
            s[f"in{i}_1"] = cur
            s[f"c{i}a"] = net._conv3(f"conv{i}_1", cur, buf(f"c{i}a", hh, ww, co), co, LK)
            if i < 5:
                s[f"p{i}"] = buf(f"p{i}", hh // 2, ww // 2, co)
                s[f"c{i}"] = net._conv3(f"conv{i}_2", s[f"c{i}a"], buf(f"c{i}", hh, ww, co), co, LK, pool_out=s[f"p{i}"])
                cur, hh, ww = s[f"p{i}"], hh // 2, ww // 2
            else:
                s[f"c{i}"] = cur = net._conv3(f"conv{i}_2", s[f"c{i}a"], buf(f"c{i}", hh, ww, co), co, LK)
        for i in range(6, 10):
            co = nf * 2 ** (9 - i)
            wu, bu = net._packed(f"upv{i}", "convT")
            s[f"u{i}"] = _conv(_lib.CONVT, cur, wu, bu, buf(f"u{i}", hh * 2, ww * 2, co), co, _lib.ACT_NONE)
            hh, ww = hh * 2, ww * 2
            s[f"c{i}a"] = net._conv3(f"conv{i}_1", s[f"u{i}"], buf(f"c{i}a", hh, ww, co), co, LK, x1=s[f"c{10 - i}"])
            s[f"c{i}"] = cur = net._conv3(f"conv{i}_2", s[f"c{i}a"], buf(f"c{i}", hh, ww, co), co, LK)
        w10, b10 = net._packed("conv10_1")
        pred = self.scr.get("pred", (n, net.out_nc, h, w), torch.float32)
        _conv(_lib.CONV1, cur, w10, b10, pred, net.out_nc, _lib.ACT_NONE, out_mode=_lib.OUT_NCHW_F32,
              resid_nchw=x if net.res else None)
        s["pred"] = pred
        return pred, s

    def backward(self, gpred, s, grad_allreduce=False):
        net, nf = self.net, self.net.nf
        self._ar = grad_allreduce
        n, _, h, w = gpred.shape
        LK = _lib.ACT_LEAKY
        self.flat_g.zero_()
        self.dw_flat.zero_()
        # 1x1 head: gradient w.r.t. conv9_2's pre-activation, dW10, db10 and conv9_2's bias gradient in one kernel
        g = self.scr.get("g_c9", (n, h, w, nf))
        m10 = net.conv10_1
        w10 = m10.weight.detach().reshape(net.out_nc, nf).to(torch.bfloat16).float()      # the forward multiplied by bf16 weights
        L.check(L.lib().pnnp_head_bwd(gpred.data_ptr(), s["c9"].data_ptr(), w10.data_ptr(),
                                      g.data_ptr(), self._grad_view("conv10_1.weight").data_ptr(),
                                      self._grad_view("conv10_1.bias").data_ptr(), self._grad_view("conv9_2.bias").data_ptr(),
                                      n, h, w, nf, net.out_nc, LK,
                                      self._stream()), "head_bwd")
        self._bucket_done("conv10_1", self._ar)
        # decoder (every data gradient that lands on an activation output carries that activation's derivative: `mask`)
        g_skip = {}
        for i in range(9, 5, -1):
            (g,) = self._conv3_bwd(f"conv{i}_2", g, [s[f"c{i}a"]], True, dx_masks=[s[f"c{i}a"]], bias_done=(i == 9))   # head kernel summed it
            g_up, g_skip[10 - i] = self._conv3_bwd(f"conv{i}_1", g, [s[f"u{i}"], s[f"c{10 - i}"]], True)
            src = s["c5"] if i == 6 else s[f"c{i - 1}"]
            g = self._convT_bwd(f"upv{i}", g_up, src)
        # encoder
        for i in range(5, 0, -1):
            if i < 5:
                ci_ = s[f"c{i}"]
                gc = self.scr.get(f"g_c{i}", tuple(ci_.shape))
                # the pool backward holds conv{i}_2's pre-activation gradient in registers: its bias gradient is summed there
                # (r02 launch list: the four read-only passes over these tensors were 80 us of the 4.2 ms step)
                fused = 256 % (ci_.shape[3] // 8) == 0 and os.environ.get("PNNP_POOL_BIAS", "1") != "0"   # 0: the separate pass (A/B)
                if fused:
                    L.check(L.lib().pnnp_maxpool_bwd_bias(g.data_ptr(), ci_.data_ptr(), g_skip[i].data_ptr(), gc.data_ptr(),
                                                          self._grad_view(f"conv{i}_2.bias").data_ptr(), ci_.shape[0], ci_.shape[1],
                                                          ci_.shape[2], ci_.shape[3], LK, self._stream()), "maxpool_bwd_bias")
                else:
                    L.check(L.lib().pnnp_maxpool_bwd(g.data_ptr(), ci_.data_ptr(), g_skip[i].data_ptr(), gc.data_ptr(),
                                                     ci_.shape[0], ci_.shape[1], ci_.shape[2], ci_.shape[3], LK, self._stream()), "maxpool_bwd")
                g = gc
                (g,) = self._conv3_bwd(f"conv{i}_2", g, [s[f"c{i}a"]], True, dx_masks=[s[f"c{i}a"]], bias_done=fused)
            else:
                (g,) = self._conv3_bwd(f"conv{i}_2", g, [s[f"c{i}a"]], True, dx_masks=[s[f"c{i}a"]])
            res = self._conv3_bwd(f"conv{i}_1", g, [s[f"in{i}_1"]], i > 1)
            if i > 1:
                g = res[0]
        if self._ar and self.comm_stream is not None:        # every bucket's all-reduce has to land before the scatter
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        tab, nd = self._unpack_tab                           # reduce buffer -> the parameters' gradient layout (one launch)
        L.check(L.lib().pnnp_strided_copy_batch(tab.data_ptr(), nd, 24, self._stream()), "unpack gradients")

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        if float(value) != self._lr:
            self._lr = float(value)
            self.adam_state[0:1].fill_(self._lr)

    def _step_body(self, lr_crops, hr, grad_allreduce):
        pred, saved = self.forward(lr_crops)
        gpred = self.scr.get("gpred", tuple(pred.shape), torch.float32)
        L.check(L.lib().pnnp_l1_loss(pred.data_ptr(), hr.data_ptr(), gpred.data_ptr(), pred.numel(), self.loss_sum.data_ptr(),
                                     self._stream()), "l1_loss")
        ddp = bool(grad_allreduce) and D.world()[1] > 1
        self.backward(gpred, saved, ddp)                                 # DDP: bucketed all-reduce(SUM) overlapped with the backward pass
        gscale = 1.0 / D.world()[1] if ddp else 1.0                      # average of the per-rank mean losses, folded into Adam
        L.check(L.lib().pnnp_adam_step_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                           self.flat_p.numel(), self.adam_state.data_ptr(), self.betas[0], self.betas[1], self.eps,
                                           gscale, self._stream()), "adam_step")
        self.refresh_packed()                                           # weights changed in place: re-pack for the next forward
        return pred

    def step(self, lr_crops, hr_crops, grad_allreduce=True):
        """One optimisation step on (noisy, clean) crops (CUDA fp32 NCHW).  Returns the loss as a 0-d CUDA tensor.

        The step is a fixed sequence of ~110 kernels on fixed buffers, so after one eager step per input shape (allocations,
        tensor maps) it is captured into a CUDA graph and replayed: the inputs are copied into static buffers first, the
        learning rate and Adam's step count are read from device memory, the gradient all-reduce is part of the graph.  With
        eight ranks sharing a host the eager step is bound by launch overhead (6.4 ms against 5.1 ms of kernels)."""
        key = (tuple(lr_crops.shape), bool(grad_allreduce))
        st = self._graphs.get(key)
        if st is None:
            st = self._graphs[key] = {"lr": torch.empty_like(lr_crops, dtype=torch.float32).contiguous(),
                                      "hr": torch.empty_like(hr_crops, dtype=torch.float32).contiguous(), "graph": None, "calls": 0}
        st["lr"].copy_(lr_crops)
        st["hr"].copy_(hr_crops)
        self.t += 1
        if st["graph"] is not None:
            st["graph"].replay()
            L.lib().pnnp_count_graph_launches(st["launches"])
        elif self.use_graph and st["calls"] >= 1:                        # second call with this shape: capture (capturing only
            torch.cuda.synchronize(self.device)                         # records the work, so the graph is replayed right after)
            l0 = _lib.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_body(st["lr"], st["hr"], grad_allreduce)
            st["launches"] = _lib.launch_count() - l0
            st["graph"] = g
            g.replay()
        else:
            self._step_body(st["lr"], st["hr"], grad_allreduce)
        st["calls"] += 1
        return (self.loss_sum / st["lr"].numel()).float()[0]
