"""Write-only and copy HBM bandwidth on this GPU with library kernels (torch fill_, zero_, copy_): the reference points for a
kernel whose traffic is mostly writes (the projection prologue: 2.43 GB written, 0.21 GB read)."""
import torch
dev = torch.device("cuda", 0)
n = 2432 * 1024 * 1024 // 2          # bf16 elements: 2.43 GB
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty(n, dtype=torch.bfloat16, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]
gb = n * 2 / 1e9
for name, fn, bytes_ in (("fill_ (write only)", lambda: a.fill_(1.0), gb), ("zero_ (memset)", lambda: a.zero_(), gb),
                         ("copy_ (read + write)", lambda: b.copy_(a), 2 * gb)):
    best, med = t(fn)
    print(f"{name:24s} {best:.4f} ms best, {med:.4f} median: {bytes_ / best * 1e3:.0f} GB/s ({bytes_:.2f} GB)")
