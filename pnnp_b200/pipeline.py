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
    def __init__(self, n, c, h, w, device, chunk=8, n_streams=3):
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
