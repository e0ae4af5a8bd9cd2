"""UNetSeeInDark / ResUnet with the reference's module API and state_dict keys
(archs/Unet.py:4-99, archs/ResUnet.py:3-88, archs/modules.py:130-197), forward executed by the
tcgen05/TMEM implicit-GEMM kernels in csrc/conv_tc.cu (NHWC bf16 activations, fp32 accumulation).

`Net(args)` takes the YAML `arch` dict (keys in_nc out_nc nf nframes res; the rest ignored), so the
trainers' `globals()[arch['name']](arch)` lookup (trainer_SID.py:17) resolves to these classes.
The parameter-holding sub-modules are ordinary nn.Conv2d / nn.ConvTranspose2d so `state_dict()`,
`load_weights(by_name=True)` (utils/utils.py:148-192) and `initialize_weights` work unchanged; they
are never *called* on the forward path — there is no eager/PyTorch fallback.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib


def initialize_weights(net):
    """archs/__init__.py:12-19 — Conv2d w,b ~ N(0, 0.02); ConvTranspose2d w ~ N(0, 0.02), bias untouched."""
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data.normal_(0.0, 0.02)
            if m.bias is not None:
                m.bias.data.normal_(0.0, 0.02)
        if isinstance(m, nn.ConvTranspose2d):
            m.weight.data.normal_(0.0, 0.02)


import os
_X_MODE = os.environ.get("PNNP_NO_XMODE") is None
_ONE_BOX = os.environ.get("PNNP_CONV3B", "1") != "0" and os.environ.get("PNNP_CONV_NOSPEC") is None and os.environ.get("PNNP_CONV_DBG") is None
_ONEBOX_MAX_COUT = int(os.environ.get("PNNP_CONV3B_MAX_COUT", "64"))
_ONEBOX_TWO_SRC = os.environ.get("PNNP_CONV3B_TWO_SRC", "0") == "1"
_ONEBOX_CIN = tuple(int(v) for v in os.environ.get("PNNP_CONV3B_CIN", "32,64").split(","))
_XMODE_MAX_COUT = int(os.environ.get("PNNP_XMODE_MAX_COUT", "32"))     # single-source layers up to this width take the x-shift-in-N mode


def _pad16(c):
    return (c + 15) // 16 * 16


class _PackedLayer:
    """bf16 weights in the kernel's [taps][rows][cin] layout + fp32 bias, rebuilt when the
    underlying parameters change (tracked by tensor._version / data_ptr)."""

    def __init__(self, module, kind, dtype=torch.bfloat16):
        self.module, self.kind, self.key, self.dtype = module, kind, None, dtype
        self.weight = self.bias = None

    def get(self, device):
        m = self.module
        key = (m.weight._version, m.weight.data_ptr(), None if m.bias is None else (m.bias._version, m.bias.data_ptr()),
               str(device))
        if key != self.key:
            w = m.weight.detach().to(device=device, dtype=torch.float32)
            if self.kind == "convT":                       # [Cin, Cout, 2, 2] -> [a*2+b][Cout][Cin]
                cin, cout = w.shape[0], w.shape[1]
                packed = w.permute(2, 3, 1, 0).reshape(4, cout, cin)
            elif self.kind == "conv3x":                    # [Cout, Cin, 3, 3] -> [ky][kx*Cout + co][Cin]  (x-shift in N)
                cout, cin = w.shape[0], w.shape[1]
                packed = w.permute(2, 3, 0, 1).reshape(3, 3 * cout, cin)
            else:                                          # [Cout, Cin, k, k] -> [ky*k+kx][Cout][Cin]
                cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
                packed = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
            rows, cin_p = _pad16(packed.shape[1]), _pad16(packed.shape[2])
            buf = torch.zeros((packed.shape[0], rows, cin_p), dtype=self.dtype, device=device)
            buf[:, :packed.shape[1], :packed.shape[2]] = packed.to(self.dtype)
            self.weight = buf.contiguous()
            self.bias = None if m.bias is None else m.bias.detach().to(device=device, dtype=torch.float32).contiguous()
            self.key = key
        return self.weight, self.bias


class _Workspace:
    """Activation buffers keyed by name, reused across calls with the same geometry."""

    def __init__(self):
        self.bufs, self.sig = {}, None

    def get(self, sig, name, shape, device, dtype=torch.bfloat16):
        if sig != self.sig:
            self.bufs, self.sig = {}, sig
        t = self.bufs.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self.bufs[name] = t
        return t


def _conv(mode, x0, w, b, out, cout, act, x1=None, resid=None, out_mode=_lib.OUT_NHWC_BF16, resid_nchw=None,
          pool_out=None, head=None, mask=None, mask_slope=0.2):
    """x0/x1: NHWC bf16 (n,h,w,c).  out: NHWC bf16 tensor, NCHW fp32 tensor, or None with `head`.
    pool_out: NHWC bf16 (n,h/2,w/2,cout) receiving the fused 2x2 max-pool.
    head = (w fp32 [co2, cout], b fp32 [co2], out fp32 NCHW): fused 1x1 conv on the activated output."""
    n, h, wd, c0 = x0.shape
    d = _lib.ConvDesc()
    d.mode, d.act, d.out_mode, d.n, d.h, d.w = mode, act, out_mode, n, h, wd
    d.io_f32 = int(x0.dtype == torch.float32)              # fp32-storage variant (tcgen05 kind::tf32): operands and outputs fp32
    if d.io_f32 and (w.dtype != torch.float32 or (x1 is not None and x1.dtype != torch.float32)):
        raise RuntimeError("pnnp_b200: fp32-storage conv needs fp32 activations and fp32 packed weights")
    d.in0, d.cin0 = x0.data_ptr(), c0
    d.in1, d.cin1 = (None, 0) if x1 is None else (x1.data_ptr(), x1.shape[3])
    d.weight, d.w_rows, d.bias = w.data_ptr(), w.shape[1], _lib.ptr(b)
    d.out, d.cout = _lib.ptr(out), cout
    d.cout_stride = cout if (out is None or out_mode != _lib.OUT_NHWC_BF16) else out.shape[3]
    d.resid, d.resid_nchw, d.pool_out = _lib.ptr(resid), _lib.ptr(resid_nchw), _lib.ptr(pool_out)
    if mask is not None:          # training dgrad: out *= act'(mask), mask = the activated output of the layer being differentiated
        d.mask, d.mask_slope = mask.data_ptr(), float(mask_slope)
    if head is not None:
        hw, hb, hout = head
        d.head_w, d.head_b, d.head_out, d.head_cout = hw.data_ptr(), hb.data_ptr(), hout.data_ptr(), hw.shape[0]
    _lib.check(_lib.lib().pnnp_conv2d_tc_ex(d, _lib.stream_ptr(x0.device)), "conv2d_tc")
    return out


def _pool(x, out):
    n, h, w, c = x.shape
    _lib.check(_lib.lib().pnnp_maxpool2x2_nhwc(x.data_ptr(), out.data_ptr(), n, h, w, c, _lib.stream_ptr(x.device)), "maxpool")
    return out


def _to_nhwc16(x, out, scale=1.0):
    n, c, h, w = x.shape
    if out.dtype == torch.float32:
        if scale != 1.0:
            raise RuntimeError("pnnp_b200: the fp32 input conversion has no scale")
        _lib.check(_lib.lib().pnnp_nchw_to_nhwc16_f32(x.data_ptr(), out.data_ptr(), n, c, h, w, _lib.stream_ptr(x.device)),
                   "nchw_to_nhwc16_f32")
        return out
    _lib.check(_lib.lib().pnnp_nchw_to_nhwc16(x.data_ptr(), out.data_ptr(), n, c, h, w, float(scale),
                                              _lib.stream_ptr(x.device)), "nchw_to_nhwc16")
    return out


def _first_conv(x, module, out, act):
    """First layer fused with the pack boundary (csrc/conv_first.cu): NCHW fp32 packed planes (<= 4 channels) -> conv3x3 + bias +
    activation -> NHWC bf16, reading the module's own fp32 parameters (rounded to bf16 in the kernel, like the packed layers)."""
    n, c, h, w = x.shape
    wt, b = module.weight.detach(), module.bias
    if wt.dtype != torch.float32 or not wt.is_contiguous():
        wt = wt.float().contiguous()
    b = None if b is None else b.detach().float().contiguous()
    _lib.check(_lib.lib().pnnp_conv_first_nchw(x.data_ptr(), wt.data_ptr(), _lib.ptr(b), out.data_ptr(), n, c, h, w, wt.shape[0], act,
                                               _lib.stream_ptr(x.device)), "conv_first_nchw")
    return out


_FUSED_FIRST = os.environ.get("PNNP_NO_FUSED_FIRST") is None


def _use_first_conv(c, cout):
    return _FUSED_FIRST and os.environ.get("PNNP_FUSED_FIRST", "1") != "0" and c <= 4 and cout % 16 == 0 and 16 <= cout <= 64


class _TCNet(nn.Module):
    def _check_input(self, x):
        _lib.require_cuda(x, "network input")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise RuntimeError("pnnp_b200: the tcgen05 forward is inference-only in this build; "
                               "wrap the call in torch.no_grad() / net.eval()")
        n, c, h, w = x.shape
        if h % 16 or w % 16:
            raise RuntimeError(f"pnnp_b200: h and w must be multiples of 16, got {h}x{w} "
                               "(the reference pads by reflection first, trainer_SID.py:221-226)")
        if c > 16:
            raise RuntimeError("pnnp_b200: at most 16 input channels")
        return x.float().contiguous()

    # Accuracy class of the forward (north_star: "within 1e-3 max-abs in fp32 (bf16 variant reported separately)"):
    #   "bf16" (default)  activations / weights bf16 in HBM, tcgen05 kind::f16, fp32 accumulation — the fast path;
    #   "tf32"            activations / weights fp32 in HBM, tcgen05 kind::tf32 (fp32 operands, 10-bit mantissa products, fp32
    #                     accumulation) — UNetSeeInDark inference; twice the bytes, for runs that need the fp32 tolerance under any init.
    # Set per network (`net.precision = "tf32"`), by the YAML arch dict (`precision: tf32`) or PNNP_UNET_PRECISION.
    precision = None

    def _dtype(self):
        p = self.precision or (self.args or {}).get("precision") or os.environ.get("PNNP_UNET_PRECISION", "bf16")
        if p not in ("bf16", "tf32"):
            raise RuntimeError(f"pnnp_b200: precision {p!r} (bf16 or tf32)")
        return torch.float32 if p == "tf32" else torch.bfloat16

    def _packed(self, name, kind="conv"):
        cache = self.__dict__.setdefault("_pack_cache", {})
        dt = self._dtype()
        key = (name, kind) if dt == torch.bfloat16 else (name, kind, "f32")
        if key not in cache:
            cache[key] = _PackedLayer(self.get_submodule(name), kind, dt)
        return cache[key].get(next(self.parameters()).device)

    def _conv3(self, name, x0, out, cout, act, **kw):
        """3x3 stride-1 conv layer `name`: the x-shift-in-N kernel mode for narrow layers (Cout <= 64, where a
        128 x Cout MMA cannot amortise its A-operand read), the per-tap mode otherwise."""
        # measured on B200 (profiles/): x-mode wins when 3*Cout <= 128 (four TMEM accumulators / epilogue groups
        # stay available) or when K is long (two-source decoder convs); otherwise the per-tap mode does.
        narrow = cout <= _XMODE_MAX_COUT or (cout <= 64 and kw.get("x1") is not None)
        # one-box mode (csrc/conv_tc.cu MODE_CONV3B, r02): single 32-channel source, cout <= _ONEBOX_MAX_COUT, plain / pool / head / residual epilogue
        if (_ONE_BOX and x0.dtype == torch.bfloat16 and (kw.get("x1") is None or _ONEBOX_TWO_SRC) and x0.shape[3] in _ONEBOX_CIN and cout % 16 == 0
                and cout <= _ONEBOX_MAX_COUT and kw.get("mask") is None and not (kw.get("pool_out") is not None and kw.get("head") is not None)
                and not (kw.get("resid") is not None and kw.get("pool_out") is not None)):
            w, b = self._packed(name)
            return _conv(_lib.CONV3B, x0, w, b, out, cout, act, **kw)
        if narrow and cout % 16 == 0 and _X_MODE:
            w, b = self._packed(name, "conv3x")
            return _conv(_lib.CONV3X, x0, w, b, out, cout, act, **kw)
        w, b = self._packed(name)
        return _conv(_lib.CONV3, x0, w, b, out, cout, act, **kw)

    def _ws(self):
        return self.__dict__.setdefault("_workspace", _Workspace())


class UNetSeeInDark(_TCNet):
    """archs/Unet.py:4-99."""

    def __init__(self, args=None):
        super().__init__()
        self.args = args
        self.nframes = args['nframes']
        self.cf = args['nframes'] // 2
        self.res = args['res']
        nf, in_nc, out_nc = args['nf'], args['in_nc'] * args['nframes'], args['out_nc']
        if nf % 16:
            raise RuntimeError("pnnp_b200: nf must be a multiple of 16 for the tensor-core path")
        self.nf, self.out_nc = nf, out_nc
        chans = [(in_nc, nf), (nf, nf * 2), (nf * 2, nf * 4), (nf * 4, nf * 8), (nf * 8, nf * 16)]
        for i, (ci, co) in enumerate(chans, start=1):
            setattr(self, f"conv{i}_1", nn.Conv2d(ci, co, kernel_size=3, stride=1, padding=1))
            setattr(self, f"conv{i}_2", nn.Conv2d(co, co, kernel_size=3, stride=1, padding=1))
            if i < 5:
                setattr(self, f"pool{i}", nn.MaxPool2d(kernel_size=2))
        for i, co in zip(range(6, 10), (nf * 8, nf * 4, nf * 2, nf)):
            setattr(self, f"upv{i}", nn.ConvTranspose2d(co * 2, co, 2, stride=2))
            setattr(self, f"conv{i}_1", nn.Conv2d(co * 2, co, kernel_size=3, stride=1, padding=1))
            setattr(self, f"conv{i}_2", nn.Conv2d(co, co, kernel_size=3, stride=1, padding=1))
        self.conv10_1 = nn.Conv2d(nf, out_nc, kernel_size=1, stride=1)
        self.relu = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        x = self._check_input(x)
        n, c, h, w = x.shape
        dev, ws, nf = x.device, self._ws(), self.nf
        dt = self._dtype()
        sig = (n, h, w, str(dev), dt)
        buf = lambda name, hh, ww, cc: ws.get(sig, name, (n, hh, ww, cc), dev, dt)
        L = _lib.ACT_LEAKY
        with torch.cuda.device(dev):
            fused_first = dt == torch.bfloat16 and _use_first_conv(c, nf)   # conv1_1 straight from the packed fp32 planes (csrc/conv_first.cu)
            cur = None if fused_first else _to_nhwc16(x, buf("x16", h, w, 16))
            skips = []
            hh, ww = h, w
            for i in range(1, 6):                                  # encoder (Unet.py:55-69)
                co = nf * 2 ** (i - 1)
                if i == 1 and fused_first:
                    t = _first_conv(x, self.conv1_1, buf("c1a", hh, ww, co), L)
                else:
                    t = self._conv3(f"conv{i}_1", cur, buf(f"c{i}a", hh, ww, co), co, L)
                if i < 5:                                          # conv + LeakyReLU + MaxPool2d(2) in one epilogue
                    pooled = buf(f"p{i}", hh // 2, ww // 2, co)
                    skips.append(self._conv3(f"conv{i}_2", t, buf(f"c{i}", hh, ww, co), co, L, pool_out=pooled))
                    cur, hh, ww = pooled, hh // 2, ww // 2
                else:
                    cur = self._conv3(f"conv{i}_2", t, buf(f"c{i}", hh, ww, co), co, L)
            for i in range(6, 10):                                 # decoder (Unet.py:71-89)
                co = nf * 2 ** (9 - i)
                skip = skips[9 - i]
                wu, bu = self._packed(f"upv{i}", "convT")
                up = _conv(_lib.CONVT, cur, wu, bu, buf(f"u{i}", hh * 2, ww * 2, co), co, _lib.ACT_NONE)
                hh, ww = hh * 2, ww * 2
                t = self._conv3(f"conv{i}_1", up, buf(f"c{i}a", hh, ww, co), co, L, x1=skip)   # cat([up, skip], 1)
                if i < 9 or co > 64 or self.out_nc > 4:
                    cur = self._conv3(f"conv{i}_2", t, buf(f"c{i}", hh, ww, co), co, L)
            out = torch.empty((n, self.out_nc, h, w), dtype=torch.float32, device=dev)
            if co <= 64 and self.out_nc <= 4:
                # conv9_2 + LeakyReLU + conv10_1 (+ x) in one kernel: conv9 never goes to HBM (Unet.py:90-98)
                m10 = self.conv10_1                      # 1x1 head kept in fp32 (it reads the fp32 accumulators)
                hw = m10.weight.detach().reshape(self.out_nc, co)
                hb = m10.bias.detach()
                if hw.dtype != torch.float32 or not hw.is_contiguous():
                    hw, hb = hw.float().contiguous(), hb.float().contiguous()
                self._conv3("conv9_2", t, None, co, L, head=(hw, hb, out), resid_nchw=x if self.res else None)
            else:
                w10, b10 = self._packed("conv10_1")
                _conv(_lib.CONV1, cur, w10, b10, out, self.out_nc, _lib.ACT_NONE, out_mode=_lib.OUT_NCHW_F32,
                      resid_nchw=x if self.res else None)
        return out


# ---- ResUnet building blocks with the reference's parameter names (archs/modules.py:130-197) ----
class conv3x3(nn.Module):
    """modules.py:130-138: stride-2 3x3 conv WITH bias; the 'relu' child hung on the nn.Conv2d never runs."""

    def __init__(self, in_nc, out_nc, stride=2, is_activate=True):
        super().__init__()
        self.conv = nn.Conv2d(in_nc, out_nc, kernel_size=3, padding=1, stride=stride)


class convWithBN(nn.Module):
    def __init__(self, in_c, out_c, kernel_size=3, padding=1, stride=1, is_activate=True, is_bn=True):
        super().__init__()
        self.conv = nn.Sequential(OrderedDict([
            ("conv", nn.Conv2d(in_c, out_c, kernel_size=kernel_size, padding=padding, stride=stride, bias=False))]))


class ResidualBlock(nn.Module):
    def __init__(self, in_c, out_c, is_activate=True):
        super().__init__()
        self.block = nn.Sequential(convWithBN(in_c, out_c, is_bn=False),
                                   convWithBN(out_c, out_c, is_activate=False, is_bn=False))
        if in_c != out_c:
            self.short_cut = nn.Sequential(convWithBN(in_c, out_c, kernel_size=1, padding=0, is_activate=False, is_bn=False))
        else:
            self.short_cut = nn.Sequential(OrderedDict([]))


class ResUnet(_TCNet):
    """archs/ResUnet.py:3-88.  ResidualBlock(is_activate=False) = conv(no bias)+ReLU -> conv(no bias),
    then `+= short_cut(x)` (identity, or a bias-free 1x1 conv when in != out); down-sampling is a
    stride-2 3x3 conv with bias and no activation."""

    def __init__(self, args=None):
        super().__init__()
        self.args = args
        self.nframes = args['nframes']
        self.cf = args['nframes'] // 2
        self.res = args['res']
        nf, in_nc, out_nc = args['nf'], args['in_nc'] * args['nframes'], args['out_nc']
        if nf % 16:
            raise RuntimeError("pnnp_b200: nf must be a multiple of 16 for the tensor-core path")
        self.nf, self.out_nc = nf, out_nc
        self.conv_in = nn.Conv2d(in_nc, nf, kernel_size=3, stride=1, padding=1)
        for i in range(1, 6):
            co = nf * 2 ** (i - 1)
            setattr(self, f"conv{i}", ResidualBlock(co, co, is_activate=False))
            if i < 5:
                setattr(self, f"pool{i}", conv3x3(co, co * 2))
        for i, co in zip(range(6, 10), (nf * 8, nf * 4, nf * 2, nf)):
            setattr(self, f"upv{i}", nn.ConvTranspose2d(co * 2, co, 2, stride=2))
            setattr(self, f"conv{i}", ResidualBlock(co * 2, co, is_activate=False))
        self.conv10 = nn.Conv2d(nf, out_nc, kernel_size=1, stride=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x, noise_map=None):
        x = self._check_input(x)
        n, c, h, w = x.shape
        dev, ws, nf = x.device, self._ws(), self.nf
        dt = self._dtype()
        sig = (n, h, w, str(dev), dt)
        buf = lambda name, hh, ww, cc: ws.get(sig, name, (n, hh, ww, cc), dev, dt)

        def block(i, src0, src1, hh, ww, co, head=None):
            """ResidualBlock i on (src0 [, src1]) -> NHWC bf16 (modules.py:176-197).  Narrow layers take the x-shift-in-N kernel
            mode like the UNet's (`_conv3`); the second conv adds the shortcut in its (specialised) epilogue, and the last block
            also applies conv10 there (`head`: the block's output then never reaches HBM)."""
            conv3 = self._conv3 if dt == torch.bfloat16 else (
                lambda name, x0, out_, cout, act, **kw: _conv(_lib.CONV3, x0, self._packed(name)[0], None, out_, cout, act, **kw))
            t = conv3(f"conv{i}.block.0.conv.conv", src0, buf(f"b{i}a", hh, ww, co), co, _lib.ACT_RELU, x1=src1)
            if src1 is None:
                shortcut = src0                                        # identity (in_c == out_c)
            else:
                wsc, _ = self._packed(f"conv{i}.short_cut.0.conv.conv")
                shortcut = _conv(_lib.CONV1, src0, wsc, None, buf(f"b{i}s", hh, ww, co), co, _lib.ACT_NONE, x1=src1)
            if head is not None:
                return conv3(f"conv{i}.block.1.conv.conv", t, None, co, _lib.ACT_NONE, resid=shortcut, head=head[:3], resid_nchw=head[3])
            return conv3(f"conv{i}.block.1.conv.conv", t, buf(f"b{i}", hh, ww, co), co, _lib.ACT_NONE, resid=shortcut)

        with torch.cuda.device(dev):
            if dt == torch.bfloat16 and _use_first_conv(c, nf):    # conv_in straight from the packed fp32 planes
                cur = _first_conv(x, self.conv_in, buf("cin", h, w, nf), _lib.ACT_RELU)
            else:
                x16 = _to_nhwc16(x, buf("x16", h, w, 16))
                wi, bi = self._packed("conv_in")
                cur = _conv(_lib.CONV3, x16, wi, bi, buf("cin", h, w, nf), nf, _lib.ACT_RELU)
            skips, hh, ww = [], h, w
            for i in range(1, 6):
                co = nf * 2 ** (i - 1)
                cur = block(i, cur, None, hh, ww, co)
                if i < 5:
                    skips.append(cur)
                    wp, bp = self._packed(f"pool{i}.conv")
                    cur = _conv(_lib.CONV3S2, cur, wp, bp, buf(f"d{i}", hh // 2, ww // 2, co * 2), co * 2, _lib.ACT_NONE)
                    hh, ww = hh // 2, ww // 2
            out = torch.empty((n, self.out_nc, h, w), dtype=torch.float32, device=dev)
            # conv10 (1x1, ResUnet.py:88) inside the last block's second conv: x-mode + residual + head epilogue (fp32 head weights)
            fuse_head = dt == torch.bfloat16 and _X_MODE and nf <= _XMODE_MAX_COUT and nf % 16 == 0 and self.out_nc <= 4 \
                and os.environ.get("PNNP_RESUNET_FUSED_HEAD", "1") != "0"
            if fuse_head:
                hw10 = self.conv10.weight.detach().reshape(self.out_nc, nf)
                hb10 = self.conv10.bias.detach()
                if hw10.dtype != torch.float32 or not hw10.is_contiguous():
                    hw10, hb10 = hw10.float().contiguous(), hb10.float().contiguous()
            for i in range(6, 10):
                co = nf * 2 ** (9 - i)
                wu, bu = self._packed(f"upv{i}", "convT")
                up = _conv(_lib.CONVT, cur, wu, bu, buf(f"u{i}", hh * 2, ww * 2, co), co, _lib.ACT_NONE)
                hh, ww = hh * 2, ww * 2
                if i == 9 and fuse_head:
                    block(i, up, skips[9 - i], hh, ww, co, head=(hw10, hb10, out, x if self.res else None))
                    return out
                cur = block(i, up, skips[9 - i], hh, ww, co)
            w10, b10 = self._packed("conv10")
            _conv(_lib.CONV1, cur, w10, b10, out, self.out_nc, _lib.ACT_NONE, out_mode=_lib.OUT_NCHW_F32,
                  resid_nchw=x if self.res else None)
        return out
