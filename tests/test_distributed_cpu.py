"""N>1 host logic on CPU (gloo, world_size 2): frame sharding and the metric all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pnnp_b200 import distributed as D


def test_shard_range_partitions_every_count():
    for n in (0, 1, 7, 30, 40, 256):
        for ws in (1, 2, 3, 4, 8):
            got = [i for r in range(ws) for i in D.shard_range(n, r, ws)]
            assert got == list(range(n))
            sizes = [len(D.shard_range(n, r, ws)) for r in range(ws)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = [20.0 + 0.5 * i for i in range(7)]           # per-frame PSNR of a 7-frame sweep
    mine = D.shard_range(len(frames), rank, world)
    p = sum(frames[i] for i in mine)
    s = sum(0.01 * frames[i] for i in mine)
    avg_p, avg_s, n = D.reduce_metric_sums(p, s, len(mine))
    out.put((rank, avg_p, avg_s, n))
    dist.destroy_process_group()


def test_metric_allreduce_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = [20.0 + 0.5 * i for i in range(7)]
    for _, avg_p, avg_s, n in res:
        assert n == 7
        assert avg_p == pytest.approx(sum(frames) / 7, abs=1e-12)
        assert avg_s == pytest.approx(0.01 * sum(frames) / 7, abs=1e-12)


def test_single_process_is_identity():
    assert D.reduce_metric_sums(6.0, 3.0, 3) == (2.0, 1.0, 3)
    assert D.world() == (0, 1)
