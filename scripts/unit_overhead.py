"""Per-unit overhead of the chunk kernel: one wave of 148 chains (37 clips x 4 heads, d_v = 256), time against the number of 64-token
chunks per chain.  Slope = chunk period, intercept = what a work unit costs besides its chunks (launch, TMEM allocation, barrier
setup, pipeline fill, initial-state load, drain, final-state store)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from bench import make_device_inputs
dev = torch.device("cuda", 0)
def t(fn, inner=20, reps=5):
    for _ in range(5): fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / inner)
    return best
rows = []
for nc in (1, 2, 4, 8, 16, 32, 64, 98):
    B, T, H, K, V = 37, 64 * nc, 4, 64, 256
    q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234, dev)
    o = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=dev); sT = torch.empty_like(S0)
    ms = t(lambda: gdkvm_b200.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 0, 0x2 | (1 << 8)))      # chunked, uncut
    rows.append((nc, ms))
    print(f"{nc:3d} chunks per chain: {ms * 1e3:8.1f} us", flush=True)
(n0, t0), (n1, t1) = rows[-3], rows[-1]
slope = (t1 - t0) / (n1 - n0)
print(f"chunk period {slope * 1e3:.2f} us = {slope * 1e-3 * 1.965e9:.0f} cycles at 1965 MHz; intercept {(t1 - slope * n1) * 1e3:.1f} us = "
      f"{(t1 - slope * n1) / slope:.1f} chunk periods")
