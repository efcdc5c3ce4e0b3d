"""One burst measurement (5 warm-up + 20 timed launches) of the memory op on B clips x frames x 49 tokens x 8 heads in a fresh process.
Usage: burst_time.py B frames flags"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from bench import make_device_inputs
B, F, flags = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
q, k, v, g, beta, S0 = make_device_inputs(B, F * 49, 8, 64, 256, 1234, dev)
o = torch.empty_like(v); sT = torch.empty_like(S0)
fn = lambda: gdkvm_b200.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 49, flags)
for _ in range(5): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
print(f"B {B} frames {F} flags {flags}: {e0.elapsed_time(e1) / 20:.4f} ms   plan {gdkvm_b200.plan_units(q, k, v, g, beta, frame_tokens=49, flags=flags)}")
