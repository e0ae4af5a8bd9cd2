"""DDP training step on two real GPUs (NCCL): initial-parameter broadcast, bucketed + overlapped gradient all-reduce, graph replay,
clean teardown.  Skipped on a single-GPU box (the CPU suite covers the host logic with gloo: tests/test_distributed_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ddp_training_step_on_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "ddp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(res.stdout)
    print(res.stderr[-2000:])
    assert res.returncode == 0, res.stdout + res.stderr[-2000:]
    assert res.stdout.count("PASS") >= 5 and "FAIL" not in res.stdout
