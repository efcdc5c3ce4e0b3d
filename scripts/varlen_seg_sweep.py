"""Packed variable-length clips (bench.py's extra.varlen_0.5 case and a wider spread): time per call against the number of time
segments per average clip (flags bits 8-11; 0 = the library's choice).  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gdkvm_b200
from bench import make_device_inputs
dev = torch.device("cuda", 0)
B, F, C, H, K, V = 64, 128, 49, 8, 64, 256
T = F * C
q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234, dev)
pk = lambda t: t.reshape(1, B * T, *t.shape[2:])
qp, kp, vp, gp, bp = pk(q), pk(k), pk(v), pk(g), pk(beta)
o2 = torch.empty(1, B * T, H, V, dtype=torch.bfloat16, device=dev)
sT2 = torch.empty(B, H, K, V, dtype=torch.float32, device=dev)
def lengths(spread, seed=4321):
    gen = torch.Generator().manual_seed(seed)
    w = 1.0 + spread * (2.0 * torch.rand(B, generator=gen) - 1.0)
    fr = torch.clamp((w / w.sum() * B * F).round().long(), min=1)
    fr[-1] += B * F - int(fr.sum())
    return fr
def t(fn, inner=150):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(inner): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / inner
# every number is a 150-call mean taken under sustained load (the GPU is at its power cap after the first second), two passes
for spread in (0.0, 0.5, 0.9):
    fr = lengths(spread)
    cu = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(fr * C, 0)]).to(dev)
    call = lambda nseg: gdkvm_b200.gdr_lkva_varlen_out(qp, kp, vp, gp, bp, cu, o2, sT2, None, S0, nseg << 8)
    t(lambda: call(0), 1000)
    for rep in range(2):
        row = [f"{nseg}: {t(lambda: call(nseg)):.3f}" for nseg in (0, 1, 2, 3, 4, 6, 8, 0)]
        print(f"spread {spread} (frames {int(fr.min())}..{int(fr.max())}):  " + "   ".join(row), flush=True)
