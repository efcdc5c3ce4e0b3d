"""The projection kernel at configs[1] geometry, twice (ncu target: scripts/gpurun: ncu -k regex:qkvgb_proj_kernel -s 1 -c 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(2)
x = torch.randn(64 * 6272, 256, generator=gen, device=dev).bfloat16()
w = (torch.randn(8 * 384 + 16, 256, generator=gen, device=dev) / 16).bfloat16()
b = 0.1 * torch.randn(8 * 384 + 16, generator=gen, device=dev)
for _ in range(3):
    gdkvm_b200.qkvgb_project(x, w, b, 8, 64, 256)
torch.cuda.synchronize()
