"""Host-buffer entry of the noise-synthesis path: pinned host crops in → pinned host noisy crops
out, with the H2D copy, the fused kernel and the D2H copy of successive chunks overlapped on
alternating CUDA streams (PCIe is full duplex, the kernel is far faster than either copy).

This is the call a DataLoader-side user makes in place of the per-crop generate_noisy_obs loop
(data_process/syn_datasets.py:326-337); bench.py times it as `e2e`."""
import torch

from . import _lib
from .noise import synthesize_batch
from .noise_params import ParamTable
from .rng import default_generator


class HostSynthPipeline:
    """zero_copy (opt-in, PNNP_E2E_ZERO_COPY=1; measured SLOWER in r02: 9 854 against 10 314 MP/s): ONE launch of the fused kernel that reads the pinned host crops and
    writes the pinned host result directly over PCIe (pinned memory is device-addressable under unified virtual addressing).  The
    same bytes cross the bus as in the chunked copy -> kernel -> copy pipeline, but both directions stream for the whole step with
    no pipeline fill or drain, and no staging buffers in HBM."""

    def __init__(self, n, c, h, w, device, chunk=8, n_streams=3, zero_copy=None):
        import os
        self.zero_copy = (os.environ.get("PNNP_E2E_ZERO_COPY", "0") == "1") if zero_copy is None else bool(zero_copy)
        self.shape = (n, c, h, w)
        self.chunk = min(chunk, n)
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self.d_in = [torch.empty((self.chunk, c, h, w), dtype=torch.float32, device=self.device) for _ in self.streams]
        self.d_out = [torch.empty_like(b) for b in self.d_in]

    def run(self, host_in, host_out, params, noise_code, chain=_lib.CHAIN_NUMPY, ori=False, clip=False,
            post_clip=None, generator=None, crop_id0=0):
        """host_in/host_out: pinned float32 CPU tensors of self.shape.  Returns host_out (valid
        after torch.cuda.synchronize() / the returned event)."""
        n = self.shape[0]
        assert tuple(host_in.shape) == self.shape and tuple(host_out.shape) == self.shape
        gen = default_generator if generator is None else generator
        seed_offset = gen.next()
        cur = torch.cuda.current_stream(self.device)
        table = ParamTable(params, self.device, torch_chain=(chain == _lib.CHAIN_TORCH))
        if self.zero_copy:
            if not (host_in.is_pinned() and host_out.is_pinned() and host_in.is_contiguous() and host_out.is_contiguous()):
                raise RuntimeError("pnnp_b200: zero-copy synthesis needs contiguous pinned host tensors")
            from .noise import noise_code_bits
            c, h, w = self.shape[1:]
            bits = noise_code_bits(noise_code) if isinstance(noise_code, str) else int(noise_code)
            if chain == _lib.CHAIN_NUMPY and getattr(table, "uniform_f64", False):
                bits |= _lib.CODE_UNIFORM_F64
            lo, hi = (-float("inf"), float("inf")) if post_clip is None else post_clip
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().pnnp_noise_synth(host_in.data_ptr(), host_out.data_ptr(), table.data_ptr(), n, c, h, w, bits, chain,
                                                       int(bool(ori)), int(bool(clip)), lo, hi, seed_offset[0], seed_offset[1], crop_id0,
                                                       _lib.stream_ptr(self.device)), "noise_synth (zero-copy)")
            return host_out
        ready = torch.cuda.Event()
        ready.record(cur)
        for i, s0 in enumerate(range(0, n, self.chunk)):
            k = i % len(self.streams)
            st = self.streams[k]
            m = min(self.chunk, n - s0)
            with torch.cuda.stream(st):
                st.wait_event(ready)
                din, dout = self.d_in[k][:m], self.d_out[k][:m]
                din.copy_(host_in[s0:s0 + m], non_blocking=True)
                synthesize_batch(din, None, noise_code, chain, ori, clip, post_clip, crop_id0=crop_id0 + s0,
                                 out=dout, table=table, table_row0=s0, seed_offset=seed_offset)
                host_out[s0:s0 + m].copy_(dout, non_blocking=True)
        for st in self.streams:
            cur.wait_stream(st)
        return host_out


class EvalPipeline:
    """Host-buffer entry of the eval path (trainer_SID.py:181-248 per frame): pinned uint16 RAW frame in ->
    P1 pack/normalise -> S3+N1-N3 noise synthesis at the frame's ratio -> [reflect-pad 4] -> network forward ->
    crop -> clamp / IlluminanceCorrect / PSNR + SSIM partial sums -> (3 + c) doubles back to pinned host memory.
    The H2D copy of frame i+1 runs on a copy stream while frame i computes."""

    def __init__(self, net, H, W, wp, bl, noise_code, ori=False, brightness_correct=False, device=None, depth=2):
        import torch.nn.functional as F  # noqa: F401
        self.net, self.H, self.W, self.wp, self.bl = net, H, W, wp, bl
        self.noise_code, self.ori, self.correct = noise_code, ori, brightness_correct
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.copy_stream = torch.cuda.Stream(self.device)
        self.d_raw = [torch.empty((H, W), dtype=torch.int16, device=self.device) for _ in range(depth)]
        self.h_sums = [torch.empty((1, 7), dtype=torch.float64).pin_memory() for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        self.slot = 0

    def submit(self, raw_host_i16, param, crop_id=0, seed_offset=(1997, 0)):
        """raw_host_i16: pinned CPU int16 tensor holding the uint16 sensor codes (H x W).  Returns the pinned
        (1, 3+c) float64 tensor that will hold the partial sums once the current stream has been synchronised."""
        import torch.nn.functional as F
        from .isp_ops import raw2bayer
        from .metrics import eval_partial_sums
        k = self.slot
        self.slot = (self.slot + 1) % len(self.d_raw)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[k])          # the previous user of this slot is done
            self.d_raw[k].copy_(raw_host_i16, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        cur.wait_event(self.ready[k])
        with torch.no_grad():
            hr = raw2bayer(self.d_raw[k], wp=self.wp, bl=self.bl, norm=True, clip=True)[None]
            self.consumed[k].record(cur)
            lr = synthesize_batch(hr, [param], self.noise_code, ori=self.ori, crop_id0=crop_id, seed_offset=seed_offset)
            if lr.shape[-1] % 16 or lr.shape[-2] % 16:
                dn = self.net(F.pad(lr, (4, 4, 4, 4), mode="reflect"))[..., 4:-4, 4:-4].contiguous()
            else:
                dn = self.net(lr)
            scale = float(param["ratio"]) if self.ori else 1.0
            sums = eval_partial_sums(dn, hr, scale, self.correct)
            self.h_sums[k].copy_(sums, non_blocking=True)
        return self.h_sums[k]


class SynthDenoisePipeline:
    """Host-buffer entry of the WHOLE hot path for a batch of crops — what one training / evaluation item of the reference
    costs on the CPU side of its DataLoader plus its network call (data_process/syn_datasets.py:296-347 -> trainer_SID.py:221-248):

        pinned uint16 RAW crops  --H2D-->  P1 raw2bayer(norm, clip)  ->  S2 + N1-N3 fused noise synthesis (+ the D1 post-clip)
        ->  UNetSeeInDark / ResUnet forward (tcgen05)  ->  clamp + PSNR / SSIM partial sums against the clean crops
        --D2H-->  (3 + c) float64 per crop in pinned host memory.

    Only sensor codes go in (2 B per raw pixel) and only metric sums come out, so the PCIe bytes per raw pixel are a quarter of
    HostSynthPipeline's float32 round trip.  The crops are processed in chunks: the H2D copy of chunk i + 1 (copy stream) overlaps the
    kernels of chunk i, and the first chunk of the NEXT batch is copied behind the last one of this batch (`next_host`).  Measured
    (r02, 64 crops of 4x512x512, one B200): chunks of 8 / 16 / 32 / 64 crops -> 7 000 / 7 390 / 7 560 / 7 610 raw MP/s end to end — the
    deep layers of the network fill the GPU better on larger batches, and with the next batch prefetched nothing is left to overlap
    inside a batch — so the default is one chunk of up to 64 crops.  (Running the metric pass of chunk i on a second stream under the
    convolutions of chunk i + 1 gained nothing: the persistent conv CTAs leave it no SM.)  bench.py times this call as `e2e` of the
    default workload."""

    def __init__(self, net, n, H, W, wp, bl, noise_code, device=None, chunk=64, post_clip=(-float("inf"), 1.0), depth=2):
        self.net, self.n, self.H, self.W, self.wp, self.bl = net, n, H, W, wp, bl
        self.noise_code, self.post_clip = noise_code, post_clip
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.chunk = min(chunk, n)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.d_raw = [torch.empty((self.chunk, H, W), dtype=torch.int16, device=self.device) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        self.h_sums = torch.empty((n, 7), dtype=torch.float64).pin_memory()
        self.slot = 0

    def run(self, host_raw_i16, params=None, table=None, generator=None, crop_id0=0, seed_offset=None, next_host=None):
        """host_raw_i16: pinned CPU int16 tensor (n, H, W) holding the uint16 sensor codes of n RAW crops; params: n reference-style
        parameter dicts (or a ParamTable).  Returns the pinned (n, 3 + c) float64 tensor of partial sums (metrics.finish_metrics turns
        a row into PSNR / SSIM), valid once the current stream has been synchronised.
        next_host: the batch the NEXT call will be given (a DataLoader hands batches over one ahead): its first chunk's H2D copy is
        issued behind this batch's last one, so it overlaps this batch's kernels instead of opening the next call."""
        from .isp_ops import raw2bayer
        from .metrics import eval_partial_sums
        n = self.n
        assert tuple(host_raw_i16.shape) == (n, self.H, self.W)
        if table is None:
            table = ParamTable(params, self.device)
        if seed_offset is None:
            seed_offset = (default_generator if generator is None else generator).next()
        cur = torch.cuda.current_stream(self.device)
        pre = getattr(self, "_prefetched", None)
        self._prefetched = None
        with torch.no_grad():
            for s0 in range(0, n, self.chunk):
                m = min(self.chunk, n - s0)
                k = self.slot
                self.slot = (self.slot + 1) % len(self.d_raw)
                if not (s0 == 0 and pre == (host_raw_i16.data_ptr(), k)):    # else: already on its way (prefetched by the previous call)
                    with torch.cuda.stream(self.copy_stream):
                        self.copy_stream.wait_event(self.consumed[k])        # the previous user of this slot has packed it
                        self.d_raw[k][:m].copy_(host_raw_i16[s0:s0 + m], non_blocking=True)
                        self.ready[k].record(self.copy_stream)
                if s0 + self.chunk >= n and next_host is not None:           # last chunk issued: start on the next batch
                    kn = self.slot
                    with torch.cuda.stream(self.copy_stream):
                        self.copy_stream.wait_event(self.consumed[kn])
                        self.d_raw[kn][:min(self.chunk, n)].copy_(next_host[0:min(self.chunk, n)], non_blocking=True)
                        self.ready[kn].record(self.copy_stream)
                    self._prefetched = (next_host.data_ptr(), kn)
                cur.wait_event(self.ready[k])
                hr = raw2bayer(self.d_raw[k][:m], wp=self.wp, bl=self.bl, norm=True, clip=True)
                self.consumed[k].record(cur)
                lr = synthesize_batch(hr, None, self.noise_code, post_clip=self.post_clip, crop_id0=crop_id0 + s0, table=table,
                                      table_row0=s0, seed_offset=seed_offset)
                dn = self.net(lr)
                self.h_sums[s0:s0 + m].copy_(eval_partial_sums(dn, hr, 1.0, False), non_blocking=True)
        return self.h_sums
