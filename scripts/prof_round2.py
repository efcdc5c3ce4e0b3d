"""One launch each of the round-2 kernels on a one-wave problem, for `ncu --set full --import-source on` (scripts/gpu_ncu2.sh)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from bench import make_device_inputs

dev = torch.device("cuda", 0)
B, T, H, K, V = 37, 32 * 64, 4, 64, 256                     # 148 chains x 32 chunks
q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1, dev)
gen = torch.Generator(device=dev).manual_seed(2)
do = torch.randn(B, T, H, V, generator=gen, device=dev).bfloat16()
dsT = torch.randn(B, H, K, V, generator=gen, device=dev)
for _ in range(2):
    o, sT, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
    torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, True)
    x = torch.randn(B * T // 4, 256, generator=gen, device=dev).bfloat16()
    w = (torch.randn(8 * 384 + 16, 256, generator=gen, device=dev) / 16).bfloat16()
    gdkvm_b200.qkvgb_project(x, w, None, 8, 64, 256)
torch.cuda.synchronize()
