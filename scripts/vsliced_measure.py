"""VERDICT round-1 lever (a), first step: "measure with the existing V = 128 instantiation, K side recomputed".  configs[1]'s value
columns split into two independent 128-column problems per (clip, head) chain (the state columns never interact), run as 16 heads
of d_v = 128: every CTA recomputes the K side (gates, K K^T, solve, T', P) for half the value columns.  One CTA per SM either way:
two co-resident CTAs do not fit the register file (16 warps x 32 x 96 registers = 49 k of 64 k per CTA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from bench import make_device_inputs
dev = torch.device("cuda", 0)
B, T, H, K, V = 64, 128 * 49, 8, 64, 256
q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234, dev)
o = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=dev); sT = torch.empty_like(S0)
dup = lambda t: t.repeat_interleave(2, dim=2).contiguous()
q2, k2, g2, b2 = dup(q), dup(k), dup(g), dup(beta)
v2 = v.reshape(B, T, H, 2, 128).reshape(B, T, 2 * H, 128).contiguous()
S2 = S0.reshape(B, H, K, 2, 128).permute(0, 1, 3, 2, 4).reshape(B, 2 * H, K, 128).contiguous()
o2 = torch.empty(B, T, 2 * H, 128, dtype=torch.bfloat16, device=dev); sT2 = torch.empty_like(S2)
def t(fn, inner=20):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(inner): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / inner
full = lambda: gdkvm_b200.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 49, 0)
half = lambda: gdkvm_b200.gdr_lkva_out(q2, k2, v2, g2, b2, o2, sT2, None, S2, 49, 0)
for rep in range(3):
    print(f"d_v = 256, 512 chains: {t(full):.4f} ms    value columns split, 1024 chains of d_v = 128: {t(half):.4f} ms", flush=True)
full(); half(); torch.cuda.synchronize()
print("same readout:", torch.equal(o.reshape(B, T, H, 2, 128), o2.reshape(B, T, H, 2, 128)), " same state:",
      torch.equal(sT.reshape(B, H, K, 2, 128).permute(0, 1, 3, 2, 4).reshape(B, 2 * H, K, 128), sT2))
