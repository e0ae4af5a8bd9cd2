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


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # flat gradient of this rank's half of a global batch: the mean of the two equals the full-batch gradient
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    scale = D.allreduce_mean_(g)
    out.put((rank, (g * scale).tolist()))
    dist.destroy_process_group()


def test_ddp_gradient_allreduce_gloo_world2():
    """T1 under DDP (SURVEY §8e Train): one all-reduce(SUM) of the flat gradient + a 1/world factor folded into Adam."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(10, dtype=torch.float32) * 1.5).tolist()
    assert all(g == want for _, g in res)
    assert D.allreduce_mean_(torch.ones(3)) == 1.0                 # single process: untouched


def test_training_epoch_gives_every_rank_the_same_number_of_steps():
    """trainer.train(): ranks wrap around the epoch's batches so the collective gradient all-reduce never deadlocks."""
    for n_batches in (1, 3, 8, 17):
        for world in (1, 2, 4, 8):
            per_rank = -(-n_batches // world)
            taken = [[(r * per_rank + j) % n_batches for j in range(per_rank)] for r in range(world)]
            assert len({len(t) for t in taken}) == 1
            assert set(i for t in taken for i in t) == set(range(n_batches))


def _sync_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pnnp_b200.train import UNetTrainStep

    class Step:                                            # the flat parameter buffer of a training step, without its CUDA parts
        refreshed = 0

        def refresh_packed(self):
            self.refreshed += 1
    st = Step()
    torch.manual_seed(100 + rank)                          # every process initialises its own weights (initialize_weights on CPU)
    st.flat_p = torch.randn(1000)
    before = st.flat_p.clone()
    UNetTrainStep.sync_parameters(st)                      # constructor path: no pack tables yet
    st._pack_tab = object()
    UNetTrainStep.sync_parameters(st)                      # checkpoint-reload path: re-packs the tensor-core weight layouts
    out.put((rank, before.tolist(), st.flat_p.tolist(), st.refreshed))
    dist.destroy_process_group()


def test_initial_parameters_are_broadcast_from_rank0_gloo_world2():
    """torchrun --mode train must start every replica from rank 0's weights, as DistributedDataParallel's
    constructor does (the reference's single-process DataParallel has one copy, base_trainer.py:115-118)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, rest) for r, *rest in (q.get(timeout=120) for _ in procs))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] != res[1][0]                          # different initial weights per process ...
    assert res[0][1] == res[1][1] == res[0][0]             # ... identical, rank 0's, after the broadcast
    assert res[0][2] == res[1][2] == 1
