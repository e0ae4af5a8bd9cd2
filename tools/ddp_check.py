"""Two-or-more-rank check of the DDP training step on real GPUs (NCCL):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ddp_check.py
1. every rank initialises its own weights (different seeds) -> UNetTrainStep broadcasts rank 0's: flat_p identical everywhere;
2. one eager step: the bucketed, overlapped all-reduce leaves SUM_r(grad_r) in flat_g on every rank == the sum of the ranks'
   local gradients (gathered from a twin step object that never reduces);
3. CUDA-graph capture + replays with the all-reduce inside the graph: parameters stay identical across ranks;
4. teardown: graphs dropped, barrier, destroy_process_group() must return (a watchdog kills the process after 90 s otherwise).
Prints one line per check on rank 0; exit code 0 = all passed."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import pnnp_b200 as P
    from pnnp_b200.train import UNetTrainStep
    arch = dict(in_nc=4, out_nc=4, nf=32, nframes=1, res=False)
    ok = True

    def say(name, passed, extra=""):
        nonlocal ok
        ok = ok and bool(passed)
        if rank == 0:
            print(f"{'PASS' if passed else 'FAIL'} {name} {extra}", flush=True)

    def same_everywhere(t):
        ref = t.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([float(torch.equal(ref, t))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    torch.manual_seed(100 + rank)                              # per-process initial weights, as torchrun gives them
    net = P.UNetSeeInDark(arch).to(dev)
    P.initialize_weights(net)
    step = UNetTrainStep(net, lr=1e-3)
    say("initial parameters broadcast from rank 0", same_everywhere(step.flat_p))

    twin_net = P.UNetSeeInDark(arch).to(dev)
    twin_net.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})
    twin = UNetTrainStep(twin_net, lr=1e-3)
    twin.use_graph = False
    g = torch.Generator(device=dev).manual_seed(7 + rank)      # different crops on every rank
    hr = torch.rand((2, 4, 64, 96), device=dev, generator=g) ** 2
    lr = (hr + 0.05 * torch.randn(hr.shape, device=dev, generator=g)).contiguous()
    step.use_graph = False
    loss = step.step(lr, hr)                                   # eager DDP step: buckets reduced on the communication stream
    twin.step(lr, hr, grad_allreduce=False)                    # same weights, same data, local gradient only
    torch.cuda.synchronize()
    local_g = twin.flat_g.clone()
    dist.all_reduce(local_g, op=dist.ReduceOp.SUM)
    err = ((step.flat_g - local_g).abs().max() / local_g.abs().max()).item()
    say("bucketed all-reduce == sum of the ranks' local gradients", err < 2e-3, f"(max rel diff {err:.2e}; fp32 atomics reorder sums)")
    say("parameters identical across ranks after the eager step", same_everywhere(step.flat_p), f"loss {float(loss):.5f}")

    step.use_graph = True
    for _ in range(6):                                         # eager, capture, replays
        step.step(lr, hr)
    torch.cuda.synchronize()
    say("parameters identical across ranks after graph-replayed DDP steps", same_everywhere(step.flat_p),
        f"(graphs: {sum(1 for s in step._graphs.values() if s['graph'] is not None)})")

    # teardown with a watchdog: r01 never returned from destroy_process_group after captured all-reduces
    done = threading.Event()

    def watchdog():
        if not done.wait(90.0):
            print(f"FAIL teardown: rank {rank} still inside destroy_process_group after 90 s", flush=True)
            os._exit(3)
    threading.Thread(target=watchdog, daemon=True).start()
    t0 = time.time()
    step._graphs.clear()
    twin._graphs.clear()
    del step, twin
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    done.set()
    if rank == 0:
        print(f"PASS teardown: destroy_process_group returned after {time.time() - t0:.1f} s", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
