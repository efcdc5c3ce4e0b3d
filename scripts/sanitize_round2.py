"""Small launches of the round-2 kernels for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_round2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from oracle.gdr_ref import make_inputs
B, T, H, V = 2, 3 * 64 + 10, 2, 256
q, k, v, g, beta, S0 = (x.cuda() for x in make_inputs(B, T, H, 64, V, seed=5, dtype=torch.bfloat16))
gen = torch.Generator(device="cuda").manual_seed(1)
do = torch.randn(B, T, H, V, generator=gen, device="cuda").bfloat16()
dsT = torch.randn(B, H, 64, V, generator=gen, device="cuda")
o, sT, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
grads = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, True, None, 2 << 8)
x = torch.randn(300, 128, generator=gen, device="cuda").bfloat16()
w = (torch.randn(2 * 384 + 4, 128, generator=gen, device="cuda") / 11).bfloat16()
out = gdkvm_b200.qkvgb_project(x, w, None, 2, 64, 256)
torch.cuda.synchronize()
print("ok", float(o.float().abs().max()), float(grads[0].float().abs().max()), float(out[0].float().abs().max()), flush=True)
