"""Prototype: the mixed plan as three launches of the batched kernel on two streams (uncut clips | first segment of the cut clips,
then their second segment through final_state -> initial_state), against one launch with the library's plan.  configs[1]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from bench import make_device_inputs
dev = torch.device("cuda", 0)
B, T, H, K, V = 64, 128 * 49, 8, 64, 256
q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234, dev)
o = torch.empty_like(v); sT = torch.empty_like(S0)
o2 = torch.empty_like(v); sT2 = torch.empty_like(S0)
NU = int(os.environ.get("UNCUT", "55")); T0 = 49 * 64
mid = torch.empty(B - NU, H, K, V, device=dev)
SEG1 = 1 << 8
s2 = torch.cuda.Stream()
def one():
    gdkvm_b200.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 49, 0)
def three(order):
    cur = torch.cuda.current_stream()
    s2.wait_stream(cur)
    def b0():
        with torch.cuda.stream(s2):
            gdkvm_b200.gdr_lkva_out(q[NU:, :T0], k[NU:, :T0], v[NU:, :T0], g[NU:, :T0], beta[NU:, :T0], o2[NU:, :T0], mid, None, S0[NU:], 0, SEG1 | 4)
    def a():
        gdkvm_b200.gdr_lkva_out(q[:NU], k[:NU], v[:NU], g[:NU], beta[:NU], o2[:NU], sT2[:NU], None, S0[:NU], 0, SEG1 | 4)
    def b1():
        with torch.cuda.stream(s2):
            gdkvm_b200.gdr_lkva_out(q[NU:, T0:], k[NU:, T0:], v[NU:, T0:], g[NU:, T0:], beta[NU:, T0:], o2[NU:, T0:], sT2[NU:], None, mid, 0, SEG1 | 4)
    for name in order: {"b0": b0, "a": a, "b1": b1}[name]()
    cur.wait_stream(s2)
def t(fn, inner=20):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(inner): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / inner
one(); three(("b0", "a", "b1")); torch.cuda.synchronize()
print("same readout / state:", torch.equal(o, o2), torch.equal(sT, sT2))
for rep in range(3):
    print(f"library plan {t(one):.4f} ms   three launches b0,a,b1 {t(lambda: three(('b0', 'a', 'b1'))):.4f}   a,b0,b1 {t(lambda: three(('a', 'b0', 'b1'))):.4f} ms", flush=True)
