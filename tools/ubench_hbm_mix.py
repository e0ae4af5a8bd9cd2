"""HBM bandwidth by read : write mix (torch kernels, CUDA events, best of 10): pure write (fill), pure read (sum), copy (1 : 1),
1 : 2 (uint16 -> float32 conversion, the pack's mix) and 1 : 4 (the first layer's mix, float32 -> four float32 outputs)."""
import torch

dev = torch.device("cuda:0")
n = 1 << 28                                   # 256 Mi elements


def best(fn, nbytes, reps=10):
    for _ in range(2):
        fn()
    t = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return nbytes / t / 1e6                   # GB/s


a = torch.empty(n, dtype=torch.float32, device=dev).normal_()
b = torch.empty_like(a)
h = torch.empty(n, dtype=torch.int16, device=dev).random_(0, 16000)
o4 = torch.empty((4, n // 4), dtype=torch.float32, device=dev)
print(f"pure write  (fill 1 GiB fp32)          {best(lambda: b.fill_(1.0), 4 * n):8.0f} GB/s")
print(f"pure read   (sum 1 GiB fp32)           {best(lambda: a.sum(), 4 * n):8.0f} GB/s")
print(f"copy 1 : 1  (fp32 -> fp32)             {best(lambda: b.copy_(a), 8 * n):8.0f} GB/s")
print(f"1 : 2       (int16 -> fp32 copy)       {best(lambda: b.copy_(h), 6 * n):8.0f} GB/s")
src = a[: n // 4]
print(f"1 : 4       (fp32 broadcast to 4 rows) {best(lambda: o4.copy_(src.unsqueeze(0).expand(4, -1)), 5 * n):8.0f} GB/s")
