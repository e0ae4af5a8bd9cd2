"""One process per GPU; the hot path shards by frame / crop with no data-path collective.
The only collective in eval is one all-reduce of [sum PSNR, sum SSIM, count] per sweep
(AverageMeter semantics: avg = sum / count, utils/utils.py:110-114)."""
import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(device_index=None):
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / MASTER_*); no-op for 1 rank."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or dist.is_initialized():
        return world()
    if torch.cuda.is_available():
        idx = int(os.environ.get("LOCAL_RANK", "0")) if device_index is None else device_index
        torch.cuda.set_device(idx)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep NCCL's banner off stdout (bench lines are parsed from it)
        dist.init_process_group("nccl", device_id=torch.device("cuda", idx))
    else:
        dist.init_process_group("gloo")
    return world()


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous block of item indices owned by `rank` (frame index -> rank, contiguous)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def reduce_metric_sums(psnr_sum, ssim_sum, count, device=None):
    """All-reduce(SUM) of three float64 scalars; returns (avg_psnr, avg_ssim, total_count)."""
    t = torch.tensor([psnr_sum, ssim_sum, float(count)], dtype=torch.float64,
                     device=device if device is not None else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    p, s, n = t.tolist()
    return (p / n if n else 0.0), (s / n if n else 0.0), int(n)


def allreduce_mean_(flat, world_size=None):
    """DDP gradient reduction on one flat buffer (SURVEY §8e Train): in-place all-reduce(SUM), returns the factor
    1 / world_size the caller folds into its optimiser step (mean of the per-rank mean losses, as DistributedDataParallel)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0
