"""Device-timed backward / training-forward / projection kernels on configs[1] (GDKVM_LIB selects another build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from bench import make_device_inputs, time_for

dev = torch.device("cuda", 0)
B, T, H, K, V = 64, 128 * 49, 8, 64, 256
q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234, dev)
gen = torch.Generator(device=dev).manual_seed(2)
do = torch.randn(B, T, H, V, generator=gen, device=dev).bfloat16()
dsT = torch.randn(B, H, K, V, generator=gen, device=dev)
o, sT, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
ms_b, n, _ = time_for(lambda: torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, True), 0.6)
ms_f, _, _ = time_for(lambda: torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0), 0.3)
x = torch.randn(B, T, 256, generator=gen, device=dev).bfloat16()
w = (torch.randn(8 * 384 + 16, 256, generator=gen, device=dev) / 16).bfloat16()
ms_p, _, _ = time_for(lambda: gdkvm_b200.qkvgb_project(x, w, None, 8, 64, 256), 0.3)
print(f"{os.environ.get('GDKVM_LIB', 'default'):60s} bwd {ms_b:7.3f} ms   train-fwd {ms_f:6.3f} ms   projection {ms_p:6.3f} ms", flush=True)
