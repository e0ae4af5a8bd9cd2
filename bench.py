#!/usr/bin/env python
"""bench.py — benchmark of the pnnp_b200 hot path (driver contract: see the task brief).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

Workloads (BASELINE.json configs; raw MP = Bayer samples = packed elements / 1e6):
  path64       DEFAULT — the metric BASELINE.json names, "noise-synth + UNet denoise": configs[1]'s 64 synthetic 4x512x512
               packed crops per GPU go through the fused SonyA7S2 'pgrq' noise synthesis AND the UNetSeeInDark forward
               (PNNP.yml arch, reference init) that consumes them; one step = both.  `roofline` = the dominant kernel (the
               tcgen05 conv kernel, tensor-core bound), `roofline_parts` = every part with its own CUDA-event time taken inside
               the timed region (synth: HBM; unet: tensor) plus configs[2]'s training step (incl. the DDP gradient all-reduce
               at N > 1) timed right after it.  e2e = pinned uint16 RAW crops in -> pack -> synthesis -> UNet -> PSNR/SSIM sums
               out (+ the eval all-reduce of the sums at N > 1).
  synth64      configs[1] alone: SonyA7S2 P-G/ELD noise synthesis 'pgrq' on 64 synthetic 4x512x512 packed crops per GPU;
               one step = one fused-kernel pass. HBM roofline.
  unet_sony    configs[0] on the GPU: UNetSeeInDark (PNNP.yml arch, reference init) eval forward on one
               synthetic 4x1424x2128 frame per GPU per step. Tensor-core roofline.
  imx686_eval  configs[3]: UNetSeeInDark eval on synthetic 4x1736x2312 frames, reflect-pad 4 -> net ->
               crop (trainer_LRID.py:224-229), one frame per GPU per step.
  sony_evaltest configs[4]: per frame pack + noise synthesis (ratio x100/x200/x300) + ResUnet forward +
               IlluminanceCorrect + PSNR/SSIM partial sums; e2e = pinned uint16 RAW frame in, metrics out.
  train_step   configs[2]: synthetic-pair training step, 8 crops of 4x512x512 per GPU (64 global on 8 GPUs): fused noise
               synthesis -> UNetSeeInDark forward -> L1 -> backward -> gradient all-reduce -> Adam; bf16 tensor cores.
Every rank processes its own crops / frames (weak scaling, no data-path collective); crop ids — and so
the Philox streams — are global.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "MP/s"
NOISE_CODE = "pgrq"
UNET_FLOP_PER_PIXEL = 92288.0          # BASELINE.md §3 (2 x MACs per raw pixel, UNetSeeInDark nf=32)
ARCH = dict(name="UNetSeeInDark", in_nc=4, out_nc=4, nf=32, nframes=1, use_dpsv=False, res=False,
            cascade=False, add=False, lock_wb=False)       # runfiles/SonyA7S2/PNNP.yml: arch

WORKLOADS = {
    "path64": dict(metric="raw megapixels/sec (noise synthesis + UNetSeeInDark denoise, 64x4x512x512 crops per GPU)", n=64, c=4,
                   h=512, w=512, bound="tensor", dtype="bf16 tcgen05 (fp32 accumulate) UNet + f64 NumPy-chain synthesis",
                   desc="path64 (BASELINE metric 'noise-synth+UNet denoise' on configs[1]'s crops): per GPU 64 crops of 4x512x512 "
                        "-> fused SonyA7S2 'pgrq' noise synthesis (numpy float64 chain, Philox4x32-10) -> UNetSeeInDark nf=32 "
                        "forward (configs[0]'s network, reference init, bf16 tcgen05)"),
    "synth64": dict(metric="raw megapixels/sec (noise synthesis, 64x4x512x512 crops per GPU)", n=64, c=4, h=512, w=512,
                    bound="hbm", dtype="f64",
                    desc="synth64 (BASELINE configs[1]): SonyA7S2 'pgrq' noise synthesis, 64 crops of 4x512x512 per GPU, "
                         "numpy (float64) chain, Philox4x32-10, 2 blocks per 4 elements"),
    "unet_sony": dict(metric="raw megapixels/sec (UNetSeeInDark eval forward, 4x1424x2128 frame per GPU)", n=1, c=4, h=1424,
                      w=2128, bound="tensor", dtype="bf16",
                      desc="unet_sony (BASELINE configs[0] on GPU): UNetSeeInDark nf=32 forward, one 4x1424x2128 frame, "
                           "bf16 tcgen05 + fp32 accumulate"),
    "imx686_eval": dict(metric="raw megapixels/sec (UNetSeeInDark eval, 4x1736x2312 frame per GPU, reflect-pad path)", n=1,
                        c=4, h=1736, w=2312, bound="tensor", dtype="bf16",
                        desc="imx686_eval (BASELINE configs[3]): reflect-pad 4 -> UNetSeeInDark -> crop on 4x1736x2312 frames"),
    "sony_evaltest": dict(metric="raw megapixels/sec (evaltest frame: noise synthesis + ResUnet forward + PSNR/SSIM, 4x1424x2128 per GPU)",
                          n=1, c=4, h=1424, w=2128, bound="tensor", dtype="bf16",
                          desc="sony_evaltest (BASELINE configs[4]): per frame pack + 'pgrq' synthesis at ratio 100/200/300 + ResUnet "
                               "forward + clamp/IlluminanceCorrect + PSNR/SSIM partial sums"),
    "train_step": dict(metric="raw megapixels/sec (synthetic-pair training step: noise synthesis + UNetSeeInDark fwd/bwd + Adam, "
                              "8 crops of 4x512x512 per GPU)", n=8, c=4, h=512, w=512, bound="tensor", dtype="bf16",
                       desc="train_step (BASELINE configs[2]): per GPU 8 crops 4x512x512 -> 'pgrq' synthesis -> UNetSeeInDark "
                            "forward + L1 + backward (bf16 tcgen05, fp32 accumulate) -> gradient all-reduce (DDP) -> Adam"),
}
RESUNET_FLOP_PER_PIXEL = 119424.0
# forward + data gradients + weight gradients; conv1_1 needs no data gradient
UNET_TRAIN_FLOP_PER_PIXEL = 3 * UNET_FLOP_PER_PIXEL - 2 * 9 * 4 * 32


# ----------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh).get(workload)
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._halt = index, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------
# CPU side.  The reference is Python and /root/reference is not on the GPU box, so the CPU legs run
#   * the UNMODIFIED reference from the git-ignored copy `oracle/_ref/` (oracle/build_ref.py, made by __graft_entry__.build() in the
#     build container; SHA-256 manifest checked before use)            -> cpu_baseline.kind == "reference"
#   * else the oracle port oracle/oracle_np.py                           -> cpu_baseline.kind == "port"
# Used ONLY as the reported baseline, never by the product path.
# ----------------------------------------------------------------------------------------------
class _RefImpl:
    """The reference's own functions behind the names the CPU legs call (same names as oracle_np)."""
    kind = "reference"
    what = "the unmodified reference (oracle/_ref: data_process/process.py generate_noisy_obs, archs/Unet.py, archs/ResUnet.py)"

    def __init__(self, ns):
        self.ns, self._mods = ns, {}
        self.sample_params = ns.process.sample_params
        self.sample_params_max = ns.process.sample_params_max

    def generate_noisy_obs(self, y, param=None, noise_code="p"):
        return self.ns.process.generate_noisy_obs(y, param=param, noise_code=noise_code)

    def _module(self, name):
        if name not in self._mods:
            arch = dict(ARCH, name=name)
            self._mods[name] = getattr(self.ns.archs, name)(arch).eval()
        return self._mods[name]

    def unet_forward(self, x, sd):
        import torch
        return torch.func.functional_call(self._module("UNetSeeInDark"), dict(sd), (x,))

    def resunet_forward(self, x, sd):
        import torch
        return torch.func.functional_call(self._module("ResUnet"), dict(sd), (x,))


_CPU_IMPL = None


def _oracle():
    """The CPU implementation the baseline legs time: the reference itself when oracle/_ref is intact, else the oracle port."""
    global _CPU_IMPL
    if _CPU_IMPL is not None:
        return _CPU_IMPL
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if os.environ.get("PNNP_BENCH_CPU_PORT", "0") != "1" and os.path.isdir(os.path.join(ref_root, "data_process")):
        try:
            import build_ref
            if build_ref.verify(ref_root):
                os.environ["PNNP_REFERENCE_ROOT"] = ref_root
                import ref_harness
                if os.path.realpath(ref_harness.REFERENCE_ROOT) == os.path.realpath(ref_root):
                    _CPU_IMPL = _RefImpl(ref_harness.load())
                    return _CPU_IMPL
        except Exception as e:                                         # a broken copy must not take the bench down
            print(f"bench: oracle/_ref unusable ({type(e).__name__}: {e}); CPU legs use the oracle port", file=sys.stderr)
    import oracle_np as O
    O.kind, O.what = "port", "the oracle port (oracle/oracle_np.py restatement of generate_noisy_obs / Unet.py / ResUnet.py)"
    _CPU_IMPL = O
    return O


def _cpu_impl_warm(_):
    _oracle()
    return 0


def _cpu_synth_one_crop(seed):
    import numpy as np
    O = _oracle()
    rs = np.random.RandomState(seed)
    y = rs.rand(4, 512, 512).astype(np.float32) ** 2
    np.random.seed(seed)
    p = O.sample_params("SonyA7S2")
    t0 = time.perf_counter()
    O.generate_noisy_obs(y, param=p, noise_code=NOISE_CODE)
    return time.perf_counter() - t0


def _cpu_unet_frames(wl, frames, threads, resunet=False):
    """reference init + torch CPU fp32 forward (oracle restatement of archs/Unet.py), `threads` threads."""
    import torch
    import torch.nn.functional as F
    O = _oracle()
    import pnnp_b200 as P
    torch.set_num_threads(threads)
    torch.manual_seed(1997)
    net = (P.ResUnet if resunet else P.UNetSeeInDark)(ARCH)
    P.initialize_weights(net)
    sd = net.state_dict()
    x = torch.rand((1, 4, wl["h"], wl["w"]))
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(frames):
            if resunet:
                import numpy as np
                np.random.seed(1)
                prm = O.sample_params_max("SonyA7S2", ratio=100, iso=1600)
                lr = torch.from_numpy(O.generate_noisy_obs(x[0].numpy(), param=prm, noise_code=NOISE_CODE))[None]
                y = O.resunet_forward(lr, sd)
            elif wl["w"] % 16:
                y = O.unet_forward(F.pad(x, (4, 4, 4, 4), mode="reflect"), sd)[..., 4:-4, 4:-4]
            else:
                y = O.unet_forward(x, sd)
    return time.perf_counter() - t0


def _cpu_train_steps(steps, threads, crops=2):
    """trainer_SID.py:93-101 on the CPU: per-crop generate_noisy_obs, fp32 autograd through the oracle UNet, Adam."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    O = _oracle()
    import pnnp_b200 as P
    torch.set_num_threads(threads)
    torch.manual_seed(1997)
    net = P.UNetSeeInDark(ARCH)
    P.initialize_weights(net)
    params = {k: v.clone().requires_grad_(True) for k, v in net.state_dict().items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    rs = np.random.RandomState(7)
    hr = rs.rand(crops, 4, 512, 512).astype(np.float32) ** 2
    t0 = time.perf_counter()
    for s in range(steps):
        np.random.seed(s)
        lr = np.stack([np.clip(O.generate_noisy_obs(hr[i], param=O.sample_params("SonyA7S2"), noise_code=NOISE_CODE), None, 1.0)
                       for i in range(crops)])
        opt.zero_grad()
        loss = F.l1_loss(O.unet_forward(torch.from_numpy(lr), params).clamp(0, 1), torch.from_numpy(hr))
        loss.backward()
        opt.step()
    return time.perf_counter() - t0, crops


def _cpu_path_crops(crops, threads):
    """The reference's CPU path for `crops` crops: per-crop generate_noisy_obs (NumPy, one thread) then the fp32 UNet forward
    of the batch on `threads` torch threads.  Returns (synthesis seconds, forward seconds)."""
    import numpy as np
    import torch
    O = _oracle()
    import pnnp_b200 as P
    torch.set_num_threads(threads)
    torch.manual_seed(1997)
    net = P.UNetSeeInDark(ARCH)
    P.initialize_weights(net)
    sd = net.state_dict()
    rs = np.random.RandomState(7)
    hr = rs.rand(crops, 4, 512, 512).astype(np.float32) ** 2
    np.random.seed(3)
    t0 = time.perf_counter()
    lr = np.stack([np.clip(O.generate_noisy_obs(hr[i], param=O.sample_params("SonyA7S2"), noise_code=NOISE_CODE), None, 1.0)
                   for i in range(crops)])
    t1 = time.perf_counter()
    with torch.no_grad():
        for i in range(crops):
            O.unet_forward(torch.from_numpy(lr[i:i + 1]), sd)
    return t1 - t0, time.perf_counter() - t1


def cpu_baseline(name, wl):
    cores = os.cpu_count() or 1
    if name == "path64":
        k = 4
        _cpu_path_crops(1, cores)                                # warm-up (thread pools, oneDNN primitives)
        ts, tu = _cpu_path_crops(k, cores)
        return {"value": k * 4 * 512 * 512 / 1e6 / (ts + tu), "unit": UNIT, "cores": cores, "kind": _oracle().kind,
                "sample": f"{k} of 64 crops (4x512x512): generate_noisy_obs one crop after the other (NumPy / SciPy are "
                          f"single-threaded: {ts / k * 1e3:.0f} ms per crop) + the UNetSeeInDark forward on torch CPU fp32 with {cores} "
                          f"threads ({tu / k * 1e3:.0f} ms per crop) — {_oracle().what}"}
    if name == "synth64":
        t_start, n, busy = time.perf_counter(), 0, 0.0
        while n < 64 and (time.perf_counter() - t_start) < 12.0:
            busy += _cpu_synth_one_crop(1000 + n)
            n += 1
        return {"value": n * 4 * 512 * 512 / 1e6 / busy, "unit": UNIT, "cores": 1, "kind": _oracle().kind,
                "sample": f"{n} of 64 crops (4x512x512, '{NOISE_CODE}'), sequential generate_noisy_obs "
                          f"(NumPy/SciPy are single-threaded) — {_oracle().what}"}
    if name == "train_step":
        dt, k = _cpu_train_steps(1, cores)
        return {"value": k * 4 * 512 * 512 / 1e6 / dt, "unit": UNIT, "cores": cores, "kind": _oracle().kind,
                "sample": f"1 step on {k} of 8 crops (4x512x512): generate_noisy_obs + torch CPU fp32 autograd of "
                          f"the UNetSeeInDark forward + Adam, {cores} threads — {_oracle().what}"}
    dt = _cpu_unet_frames(wl, 1, cores, resunet=(name == "sony_evaltest"))
    what = "generate_noisy_obs + resunet_forward (metrics excluded)" if name == "sony_evaltest" else "unet_forward"
    return {"value": wl["c"] * wl["h"] * wl["w"] / 1e6 / dt, "unit": UNIT, "cores": cores, "kind": _oracle().kind,
            "sample": f"1 frame 4x{wl['h']}x{wl['w']}, torch CPU fp32, {cores} threads, {what} — {_oracle().what}"}


def run_reference_arm(args, name, wl):
    """--impl reference: the reference's CPU algorithm with all the host threads it can use."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    if name == "path64":
        # all host cores on both halves: synthesis of `workers` crops in `workers` processes (NumPy is single-threaded, this is the
        # reference's DataLoader(num_workers) arrangement), then their fp32 UNet forwards on `cores` torch threads
        import multiprocessing as mp
        workers = max(1, min(cores, 16))
        pool = mp.get_context("fork").Pool(workers)                # forked before torch starts its thread pools
        pool.map(_cpu_impl_warm, range(workers))                   # every worker imports the CPU implementation outside the timed region
        import torch
        O = _oracle()
        import pnnp_b200 as P
        torch.set_num_threads(cores)
        torch.manual_seed(1997)
        net = P.UNetSeeInDark(ARCH)
        P.initialize_weights(net)
        sd = net.state_dict()
        x = torch.rand((1, 4, 512, 512))

        def one_step(seed0):
            pool.map(_cpu_synth_one_crop, [seed0 + i for i in range(workers)])
            with torch.no_grad():
                for _ in range(workers):
                    O.unet_forward(x, sd)
        for w in range(min(args.warmup, 1)):
            one_step(w * workers)
        steps = min(args.steps, 4)
        t0 = time.perf_counter()
        for s_ in range(steps):
            one_step(5000 + s_ * workers)
        dt = (time.perf_counter() - t0) * args.steps / steps
        pool.close()
        mp_step = workers * 4 * 512 * 512 / 1e6
        sample = (f"{workers} of 64 crops per step ({steps} steps timed, scaled to {args.steps}): generate_noisy_obs on "
                  f"{workers} worker processes, then the UNetSeeInDark forward per crop on torch CPU fp32 with {cores} threads")
        used = cores
    elif name == "synth64":
        import multiprocessing as mp
        workers = max(1, min(cores, 64))
        per_step = workers                                     # bounded sample: one crop per worker per step
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_cpu_impl_warm, range(workers))
            for w in range(args.warmup):
                pool.map(_cpu_synth_one_crop, [w * per_step + i for i in range(per_step)])
            t0 = time.perf_counter()
            for s in range(args.steps):
                pool.map(_cpu_synth_one_crop, [5000 + s * per_step + i for i in range(per_step)])
            dt = time.perf_counter() - t0
        mp_step = per_step * 4 * 512 * 512 / 1e6
        sample = f"{per_step} crops per step on {workers} worker processes (generate_noisy_obs)"
        used = workers
    elif name == "train_step":
        steps = min(args.steps, 2)
        dt, k = _cpu_train_steps(steps, cores)
        dt = dt * args.steps / steps
        mp_step = k * 4 * 512 * 512 / 1e6
        sample = f"{k} of 8 crops per step ({steps} steps timed, scaled to {args.steps}), torch CPU fp32 autograd on {cores} threads"
        used = cores
    else:
        steps = min(args.steps, 3)                             # ~5 s per frame on 8 cores
        rn = name == "sony_evaltest"
        _cpu_unet_frames(wl, min(args.warmup, 1), cores, resunet=rn)
        dt = _cpu_unet_frames(wl, steps, cores, resunet=rn) * args.steps / steps
        mp_step = wl["c"] * wl["h"] * wl["w"] / 1e6
        sample = f"1 frame per step ({steps} timed, scaled to {args.steps}), torch CPU fp32 on {cores} threads"
        used = cores
    value = mp_step * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if name == "synth64" else ("f64 synthesis + f32 UNet" if name == "path64" else "f32"),
        "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": _oracle().kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args, name, wl):
    import numpy as np
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    import pnnp_b200 as P
    from pnnp_b200 import _lib
    from pnnp_b200.pipeline import HostSynthPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    try:        # keep this rank (and the pinned buffers it first-touches) on the CPU cores / NUMA node next to its GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created; stdout carries the one JSON line,
        # so file descriptor 1 points at stderr while the process group and its first collective are set up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, c, h, w = wl["n"], wl["c"], wl["h"], wl["w"]
    elems = n * c * h * w
    gen = P.PhiloxGenerator(1997)
    g = torch.Generator(device=device).manual_seed(1997 + rank)
    marks = []                                                                 # path64: per-step (start, after synthesis, after UNet) events
    if name == "path64":
        torch.manual_seed(1997)
        net = P.UNetSeeInDark(ARCH).to(device).eval()
        P.initialize_weights(net)
        clean = torch.rand((n, c, h, w), device=device, generator=g) ** 2      # dark-scene distribution
        np.random.seed(1997 + rank)
        params = [P.sample_params("SonyA7S2") for _ in range(n)]
        table = P.ParamTable(params, device)
        noisy = torch.empty_like(clean)
        crop0 = rank * n

        def step():
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            P.synthesize_batch(clean, None, NOISE_CODE, _lib.CHAIN_NUMPY, post_clip=(-float("inf"), 1.0),
                               generator=gen, crop_id0=crop0, out=noisy, table=table)
            ev[1].record()
            with torch.no_grad():
                net(noisy)
            ev[2].record()
            marks.append(ev)
        algo = UNET_FLOP_PER_PIXEL * elems                                     # the dominant kernel's work: the UNet's FLOPs
        l2_note = "no flush: 268 MB of crops in, 268 MB noisy out, 1.07 GB per full-resolution activation tensor: all exceed the 126 MB L2"
        kernel = "conv_gemm_tc_kernel (the 23 conv launches of the UNet forward; share of the step in profiles/)"
    elif name == "synth64":
        clean = torch.rand((n, c, h, w), device=device, generator=g) ** 2      # dark-scene distribution
        np.random.seed(1997 + rank)
        params = [P.sample_params("SonyA7S2") for _ in range(n)]
        table = P.ParamTable(params, device)
        out = torch.empty_like(clean)
        crop0 = rank * n

        def step():
            P.synthesize_batch(clean, None, NOISE_CODE, _lib.CHAIN_NUMPY, post_clip=(-float("inf"), 1.0),
                               generator=gen, crop_id0=crop0, out=out, table=table)
        algo = elems * 8.0                                                     # 4 B read + 4 B write per element
        l2_note = "no flush: 268 MB in + 268 MB out per step exceed the 126 MB L2"
        kernel = "noise_synth_fast_kernel"
    elif name == "train_step":
        from pnnp_b200.train import UNetTrainStep
        torch.manual_seed(1997)
        net = P.UNetSeeInDark(ARCH).to(device)
        P.initialize_weights(net)
        trainer = UNetTrainStep(net, lr=1e-4)
        clean = torch.rand((n, c, h, w), device=device, generator=g) ** 2     # (CUDA-graph replay at every N: 8 ranks launching eagerly
                                                                              # from one host take 6.4 ms per step, replayed 5.2)
        np.random.seed(1997 + rank)
        params = [P.sample_params("SonyA7S2") for _ in range(n)]
        table = P.ParamTable(params, device)
        noisy = torch.empty_like(clean)
        crop0 = rank * n
        losses = []

        def step():
            P.synthesize_batch(clean, None, NOISE_CODE, _lib.CHAIN_NUMPY, post_clip=(-float("inf"), 1.0),
                               generator=gen, crop_id0=crop0, out=noisy, table=table)
            losses.append(trainer.step(noisy, clean))
        algo = UNET_TRAIN_FLOP_PER_PIXEL * elems
        l2_note = "no flush: saved activations of a step (2.7 GB) exceed the 126 MB L2"
        kernel = "conv_gemm_tc_kernel (fwd + dgrad) + wgrad_tc_kernel"
    elif name == "sony_evaltest":
        from pnnp_b200.pipeline import EvalPipeline
        from pnnp_b200.metrics import eval_partial_sums
        torch.manual_seed(1997)
        net = P.ResUnet(dict(ARCH, name="ResUnet")).to(device).eval()
        P.initialize_weights(net)
        raw = (512 + (torch.rand((2 * h, 2 * w), device=device, generator=g) ** 2) * (16383 - 512)).to(torch.int16)
        np.random.seed(1997 + rank)
        sweep = [P.sample_params_max("SonyA7S2", ratio=r, iso=1600) for r in (100, 200, 300)]
        state = {"i": 0}

        def step():
            with torch.no_grad():
                prm = sweep[state["i"] % 3]
                state["i"] += 1
                hr = P.raw2bayer(raw, wp=16383, bl=512, norm=True, clip=True)[None]
                lr = P.synthesize_batch(hr, [prm], NOISE_CODE, crop_id0=rank, seed_offset=(1997, state["i"]))
                eval_partial_sums(net(lr), hr, 1.0, True)
        algo = RESUNET_FLOP_PER_PIXEL * n * c * h * w
        l2_note = "no flush: level-1/2 activations (194 MB per tensor) exceed the 126 MB L2"
        kernel = "conv_gemm_tc_kernel (ResUnet convs) + noise_synth_kernel + ssim_mse_kernel"
    else:
        torch.manual_seed(1997)
        net = P.UNetSeeInDark(ARCH).to(device).eval()
        P.initialize_weights(net)
        frame = torch.rand((n, c, h, w), device=device, generator=g)
        pad = (w % 16) != 0

        def forward(x):
            with torch.no_grad():
                if pad:                                                        # trainer_LRID.py:224-229
                    return net(F.pad(x, (4, 4, 4, 4), mode="reflect"))[..., 4:-4, 4:-4]
                return net(x)

        def step():
            forward(frame)
        hp, wp_ = (h + 8, w + 8) if pad else (h, w)
        algo = UNET_FLOP_PER_PIXEL * n * c * hp * wp_                          # FLOPs on the padded extent (per RAW pixel)
        l2_note = "no flush: level-1/2 activations (194 MB per tensor) exceed the 126 MB L2"
        kernel = "conv_gemm_tc_kernel (all conv layers; share of step in profiles/)"

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    marks.clear()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - l0
    ms_step = allmax(e0.elapsed_time(e1)) / args.steps
    value = world * elems / 1e6 / (ms_step / 1e3)

    # ---- end-to-end through the public host-buffer API (pinned host in, host result out)
    e2e_steps = max(3, min(args.steps, 20))
    parts = None
    if name == "path64":
        # per-part times taken INSIDE the timed region above (events between the synthesis launch and the UNet's launches)
        synth_ms = allmax(sum(a.elapsed_time(b) for a, b, _ in marks) / len(marks))
        unet_ms = allmax(sum(b.elapsed_time(c_) for _, b, c_ in marks) / len(marks))
        parts = {"synth_ms": synth_ms, "unet_ms": unet_ms}
        from pnnp_b200.pipeline import SynthDenoisePipeline
        from pnnp_b200 import distributed as D
        from pnnp_b200.metrics import finish_metrics
        raw_dev = (512 + clean.reshape(n, c, h, w) * (16383 - 512)).round().clamp(0, 16383)       # sensor codes of the same crops
        raw_host = torch.empty((n, 2 * h, 2 * w), dtype=torch.int16).pin_memory()
        mosaic = torch.empty((n, 2 * h, 2 * w), device=device)
        mosaic[:, 0::2, 0::2], mosaic[:, 0::2, 1::2] = raw_dev[:, 0], raw_dev[:, 1]              # R G1 / G2 B (isp_ops.py:87-90)
        mosaic[:, 1::2, 1::2], mosaic[:, 1::2, 0::2] = raw_dev[:, 2], raw_dev[:, 3]
        raw_host.copy_(mosaic.to(torch.int16).cpu())
        del mosaic, raw_dev
        pipe = SynthDenoisePipeline(net, n, 2 * h, 2 * w, 16383, 512, NOISE_CODE, device, chunk=int(os.environ.get("PNNP_E2E_CHUNK", "64")))
        e2e_metrics = {}

        def e2e_step():
            sums = pipe.run(raw_host, table=table, generator=gen, crop_id0=crop0, next_host=raw_host)      # batches arrive one ahead
            torch.cuda.current_stream().synchronize()            # the metric sums are on the host: the step's result
            rows = finish_metrics(sums, c, h, w)
            ps, ss = sum(r["PSNR"] for r in rows), sum(r["SSIM"] for r in rows)
            e2e_metrics["psnr"], e2e_metrics["ssim"], e2e_metrics["crops"] = D.reduce_metric_sums(ps, ss, n, device)   # eval all-reduce
        h2d, d2h = n * 2 * h * 2 * w * 2, n * 7 * 8
        api = ("pnnp_b200.pipeline.SynthDenoisePipeline.run: pinned uint16 RAW crops in -> pack -> synthesis -> UNetSeeInDark -> "
               "PSNR/SSIM partial sums out (+ one all-reduce of [sum PSNR, sum SSIM, count] per step at N > 1)")
    elif name == "synth64":
        pipe = HostSynthPipeline(n, c, h, w, device, chunk=int(os.environ.get("PNNP_E2E_CHUNK", "8")),
                                 n_streams=int(os.environ.get("PNNP_E2E_STREAMS", "3")))
        host_in = torch.empty((n, c, h, w), dtype=torch.float32).pin_memory()
        host_in.copy_(clean.cpu())
        host_out = torch.empty_like(host_in).pin_memory()

        def e2e_step():
            pipe.run(host_in, host_out, params, NOISE_CODE, generator=gen, crop_id0=crop0, post_clip=(-float("inf"), 1.0))
        h2d, d2h = elems * 4 + n * 128, elems * 4
        api = "pnnp_b200.pipeline.HostSynthPipeline.run (pinned host crops in, pinned host noisy crops out)"
    elif name == "train_step":
        # the loop body of trainer_SID.py:93-101 with host-resident clean crops (what the DataLoader hands over):
        # H2D of the clean crops, synthesis + step on the device, the loss back on the host every step
        host_clean = [clean.cpu().pin_memory(), clean.cpu().pin_memory()]
        dev_clean = torch.empty_like(clean)
        host_loss = torch.zeros(1).pin_memory()
        st3 = {"i": 0}

        def e2e_step():
            i = st3["i"]
            st3["i"] += 1
            dev_clean.copy_(host_clean[i % 2], non_blocking=True)
            P.synthesize_batch(dev_clean, None, NOISE_CODE, _lib.CHAIN_NUMPY, post_clip=(-float("inf"), 1.0),
                               generator=gen, crop_id0=crop0, out=noisy, table=table)
            host_loss.copy_(trainer.step(noisy, dev_clean).reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the reference logs loss.item() every step
        h2d, d2h = elems * 4, 4
        api = "pnnp_b200.train.UNetTrainStep.step on pinned host clean crops: H2D + synthesis + fwd/bwd/Adam + loss D2H"
    elif name == "sony_evaltest":
        pipe = EvalPipeline(net, 2 * h, 2 * w, 16383, 512, NOISE_CODE, brightness_correct=True, device=device)
        host_raw = [raw.cpu().pin_memory(), raw.cpu().pin_memory()]
        st2 = {"i": 0}

        def e2e_step():
            i = st2["i"]
            st2["i"] += 1
            pipe.submit(host_raw[i % 2], sweep[i % 3], crop_id=rank, seed_offset=(1997, i))
        h2d, d2h = 2 * h * 2 * w * 2 + 128, 7 * 8
        api = "pnnp_b200.pipeline.EvalPipeline.submit: pinned uint16 RAW frame in, PSNR/SSIM partial sums out"
    else:
        host_in = torch.empty((n, c, h, w), dtype=torch.float32).pin_memory()
        host_in.copy_(frame.cpu())
        host_out = torch.empty_like(host_in).pin_memory()
        dev_in = torch.empty_like(frame)

        def e2e_step():
            dev_in.copy_(host_in, non_blocking=True)
            host_out.copy_(forward(dev_in), non_blocking=True)
        h2d, d2h = elems * 4, elems * 4
        api = "UNetSeeInDark.forward on a pinned host frame: H2D + forward + D2H of the denoised frame"
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_value = world * elems / 1e6 / (allmax(time.perf_counter() - t0) / e2e_steps)

    train_part = None
    if name == "path64" and not args.no_parts:
        # configs[2] right behind it: synthetic-pair training step on 8 of the crops per GPU (64 global at N = 8), CUDA-graph replay,
        # the DDP gradient all-reduce inside the step at N > 1
        from pnnp_b200.train import UNetTrainStep
        torch.manual_seed(1997)
        tnet = P.UNetSeeInDark(ARCH).to(device)
        P.initialize_weights(tnet)
        trainer = UNetTrainStep(tnet, lr=1e-4)
        tc_, tn_ = clean[:8].contiguous(), torch.empty((8, c, h, w), device=device)
        t_losses = []

        def train_step():
            P.synthesize_batch(tc_, None, NOISE_CODE, _lib.CHAIN_NUMPY, post_clip=(-float("inf"), 1.0), generator=gen,
                               crop_id0=crop0, out=tn_, table=table)
            t_losses.append(trainer.step(tn_, tc_))
        for _ in range(4):                                       # eager step, graph capture, replays
            train_step()
        barrier()
        t_steps = max(5, min(args.steps, 20))
        te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        te0.record()
        for _ in range(t_steps):
            train_step()
        te1.record()
        barrier()
        train_part = {"ms_per_step": allmax(te0.elapsed_time(te1)) / t_steps, "steps": t_steps, "crops_per_gpu": 8,
                      "loss_first_last": [float(t_losses[0]), float(t_losses[-1])]}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        burst, sustained = float(peaks["bf16_tflops"]), float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        if args.precision == "tf32":                            # dense tf32 = half the bf16 rate (1.1 vs 2.25 PFLOP/s nominal)
            burst, sustained = burst / 2, sustained / 2
        if name == "path64":
            # the dominant kernel's own time inside the timed region; burst peak for a short timed region, sustained for a long one
            timed_s = ms_step * args.steps / 1e3
            peak, which = (sustained, "sustained cuBLAS bf16 (timed region >= 1 s)") if timed_s >= 1.0 else \
                          (burst, f"burst cuBLAS bf16 (timed region {timed_s:.2f} s < 1 s)")
            achieved, unit = algo / (parts["unet_ms"] / 1e3) / 1e12, "TFLOP/s"
        elif wl["bound"] == "hbm":
            achieved, peak, unit = algo / (ms_step / 1e3) / 1e9, float(peaks["hbm_gbs"]), "GB/s"
            which = "burst copy"
        else:
            achieved, unit = algo / (ms_step / 1e3) / 1e12, "TFLOP/s"
            timed_s = ms_step * args.steps / 1e3
            peak, which = (sustained, "sustained cuBLAS bf16 (timed region >= 1 s)") if timed_s >= 1.0 else \
                          (burst, f"burst cuBLAS bf16 (timed region {timed_s:.2f} s < 1 s)")
        line = {
            "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": wl["desc"], "per_gpu_shape": [n, c, h, w], "l2": l2_note},
            "roofline": {"bound": wl["bound"], "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                         "traffic": ncu_traffic(name), "peak_source": f"{peak_src}, {which}", "kernel": kernel,
                         "algorithmic_work_per_step": algo},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": api,
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if unit == "TFLOP/s":
            line["roofline"].update(frac_of_burst=achieved / burst, frac_of_sustained=achieved / sustained)
        if name == "path64":
            hbm = float(peaks["hbm_gbs"])
            sy = elems * 8.0 / (parts["synth_ms"] / 1e3) / 1e9
            tr = ncu_traffic("path64_parts") or {}
            line["roofline"].update(frac_of_burst=achieved / burst, frac_of_sustained=achieved / sustained,
                                    kernel_ms_per_step=parts["unet_ms"], traffic=tr.get("unet"))
            line["roofline_parts"] = {
                "synth": {"bound": "hbm", "kernel": "noise_synth_fast_kernel", "ms_per_step": parts["synth_ms"], "achieved": sy,
                          "peak": hbm, "unit": "GB/s", "frac": sy / hbm, "algorithmic_work_per_step": elems * 8.0,
                          "traffic": tr.get("synth")},
                "unet": {"bound": "tensor", "kernel": "conv_gemm_tc_kernel x 23", "ms_per_step": parts["unet_ms"], "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "frac_of_burst": achieved / burst,
                         "frac_of_sustained": achieved / sustained, "algorithmic_work_per_step": algo, "traffic": tr.get("unet")},
            }
            if train_part:
                ta = UNET_TRAIN_FLOP_PER_PIXEL * 8 * c * h * w / (train_part["ms_per_step"] / 1e3) / 1e12
                line["roofline_parts"]["train_step"] = dict(
                    train_part, bound="tensor", kernel="conv_gemm_tc_kernel (fwd + dgrad) + wgrad_nhwc_kernel", achieved=ta, peak=burst,
                    unit="TFLOP/s", frac=ta / burst, frac_of_sustained=ta / sustained,
                    value_mp_s=world * 8 * c * h * w / 1e6 / (train_part["ms_per_step"] / 1e3),
                    note="BASELINE configs[2]: synthesis + UNet fwd/bwd + Adam, CUDA-graph replay"
                         + (f", DDP gradient all-reduce over {world} ranks inside the step" if world > 1 else ""))
            line["e2e"]["eval_allreduce"] = {"ranks": world, "avg_psnr": e2e_metrics.get("psnr"), "avg_ssim": e2e_metrics.get("ssim"),
                                             "crops": e2e_metrics.get("crops")}
        if name == "train_step":
            line["config"]["global_batch"] = world * n
            line["loss_first_last"] = [float(losses[0]), float(losses[-1])]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(name, wl)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown.  Graphs that captured all-reduces go first (r01: destroy_process_group() never returned with them alive; r02,
        # tools/ddp_check.py: 0.5 s once they are dropped); a watchdog still ends the process if NCCL hangs, after the JSON line.
        sys.stdout.flush()
        sys.stderr.flush()
        if "trainer" in locals():
            trainer.close()
        torch.cuda.synchronize()
        watchdog = threading.Timer(60.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        watchdog.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="pnnp_b200", choices=["pnnp_b200", "reference"])
    ap.add_argument("--workload", default="path64", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parts", action="store_true", help="path64: skip the training-step part")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"],
                    help="network accuracy class: bf16 storage + kind::f16 MMAs (default) or fp32 storage + kind::tf32 MMAs")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.precision == "tf32":
        os.environ["PNNP_UNET_PRECISION"] = "tf32"
        wl = dict(wl, dtype="tf32 tcgen05 (fp32 storage, fp32 accumulate)" + (" UNet + f64 NumPy-chain synthesis" if args.workload == "path64" else ""),
                  desc=wl["desc"] + " [fp32-storage / kind::tf32 variant: the roofline peak is half the measured bf16 peak]")
    if args.impl == "reference":
        return run_reference_arm(args, args.workload, wl)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
                                   "29533", os.path.abspath(__file__)] + sys.argv[1:])
    run_gpu_arm(args, args.workload, wl)


if __name__ == "__main__":
    main()
