#!/usr/bin/env python
"""bench.py — headline benchmark of the pnnp_b200 hot path (driver contract: see the task brief).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

Default workload (`synth64`) is BASELINE.json configs[1]: SonyA7S2 P-G/ELD noise synthesis
(Poisson + Tukey-lambda + row + quantisation, noise_code 'pgrq') on 64 synthetic 4x512x512 packed
crops per GPU.  A "step" is one fused-kernel pass over that batch.  Metric: raw megapixels / s
(raw MP = packed elements / 1e6), whole-job aggregate over all ranks (weak scaling: every rank
synthesises its own 64 crops; crop ids — and therefore Philox streams — are global).

One JSON line on stdout (rank 0).  Keys beyond the base contract: roofline, cpu_baseline, e2e,
gpu_launches, clocks.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "raw megapixels/sec (noise synthesis, 64x4x512x512 crops per GPU)"
UNIT = "MP/s"
N_CROPS, C, HW = 64, 4, 512
ELEMS = N_CROPS * C * HW * HW                      # 67 108 864 packed elements per GPU per step
ALGO_BYTES = ELEMS * 8                             # 4 B read + 4 B write per element
NOISE_CODE = "pgrq"


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary, if any."""
    p = os.path.join(ROOT, "profiles", "noise_synth_traffic.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh).get("dram_bytes_per_launch")
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
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
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


def synth_inputs(torch, device, rank):
    """Synthetic dark-scene crops (u^2) and 64 parameter rows, as SURVEY §8d config 2 prescribes."""
    import numpy as np
    import pnnp_b200 as P
    g = torch.Generator(device=device).manual_seed(1997 + rank)
    clean = torch.rand((N_CROPS, C, HW, HW), device=device, generator=g) ** 2
    np.random.seed(1997 + rank)
    params = [P.sample_params("SonyA7S2") for _ in range(N_CROPS)]
    return clean, params


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's generate_noisy_obs: same NumPy/SciPy calls)
# ----------------------------------------------------------------------------------------------
def _cpu_one_crop(seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_np as O
    rs = np.random.RandomState(seed)
    y = rs.rand(C, HW, HW).astype(np.float32) ** 2
    np.random.seed(seed)
    p = O.sample_params("SonyA7S2")
    t0 = time.perf_counter()
    O.generate_noisy_obs(y, param=p, noise_code=NOISE_CODE)
    return time.perf_counter() - t0


def cpu_baseline_single(budget_s=12.0, max_crops=N_CROPS):
    """One thread, sequential crops (NumPy/SciPy are single-threaded here), bounded by time."""
    t_start, n, busy = time.perf_counter(), 0, 0.0
    while n < max_crops and (time.perf_counter() - t_start) < budget_s:
        busy += _cpu_one_crop(1000 + n)
        n += 1
    mp = n * C * HW * HW / 1e6
    return {"value": mp / busy, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} of {N_CROPS} crops (4x512x512, '{NOISE_CODE}'), sequential, oracle/oracle_np.py::generate_noisy_obs"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port — the reference itself is
    Python and /root/reference does not exist on the GPU box) on all host cores, one worker
    process per core like the reference's DataLoader workers."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_step = min(N_CROPS, workers)                       # bounded sample: one crop per worker per step
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for w in range(args.warmup):
            pool.map(_cpu_one_crop, [w * per_step + i for i in range(per_step)])
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_cpu_one_crop, [5000 + s * per_step + i for i in range(per_step)])
        dt = time.perf_counter() - t0
    mp_per_step = per_step * C * HW * HW / 1e6
    value = mp_per_step * args.steps / dt
    sample = f"{per_step} crops per step on {workers} worker processes (oracle port of generate_noisy_obs)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": "synth64: SonyA7S2 'pgrq' noise synthesis, 4x512x512 crops",
                                            "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
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
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clean, params = synth_inputs(torch, device, rank)
    table = P.ParamTable(params, device)
    out = torch.empty_like(clean)
    gen = P.PhiloxGenerator(1997)
    crop0 = rank * N_CROPS

    def step():
        P.synthesize_batch(clean, None, NOISE_CODE, _lib.CHAIN_NUMPY, post_clip=(-float("inf"), 1.0),
                           generator=gen, crop_id0=crop0, out=out, table=table)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
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
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * ELEMS / 1e6 / (ms_step / 1e3)

    # ---- end-to-end through the public host-buffer API (pinned host in, pinned host out)
    pipe = HostSynthPipeline(N_CROPS, C, HW, HW, device)
    host_in = torch.empty((N_CROPS, C, HW, HW), dtype=torch.float32).pin_memory()
    host_in.copy_(clean.cpu())
    host_out = torch.empty_like(host_in).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        pipe.run(host_in, host_out, params, NOISE_CODE, generator=gen, crop_id0=crop0, post_clip=(-float("inf"), 1.0))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.run(host_in, host_out, params, NOISE_CODE, generator=gen, crop_id0=crop0, post_clip=(-float("inf"), 1.0))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * ELEMS / 1e6 / (float(te.item()) / e2e_steps)

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = ALGO_BYTES / (ms_step / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synth64 (BASELINE configs[1]): SonyA7S2 'pgrq' noise synthesis, 64 crops of "
                                   "4x512x512 per GPU, numpy (float64) chain, Philox4x32-10",
                       "crops_per_gpu": N_CROPS, "crop": [C, HW, HW], "noise_code": NOISE_CODE,
                       "l2": "no flush: 268 MB in + 268 MB out per step exceed the 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": "noise_synth_kernel",
                         "algorithmic_bytes_per_launch": ALGO_BYTES},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ELEMS * 4 + N_CROPS * 128,
                    "d2h_bytes_per_step": ELEMS * 4, "api": "pnnp_b200.pipeline.HostSynthPipeline.run (pinned host in/out)",
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="pnnp_b200", choices=["pnnp_b200", "reference"])
    ap.add_argument("--workload", default="synth64")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
                                   "29533", os.path.abspath(__file__)] + sys.argv[1:])
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
