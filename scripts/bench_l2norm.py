"""Device-timed roofline of the q/k L2-norm prologue kernel at the configs[1] shape (GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
B, T, H, D = 64, 128 * 49, 8, 64
x = torch.randn(B, T, H, D, device="cuda").bfloat16()
for _ in range(3):
    y = gdkvm_b200.l2norm(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    y = gdkvm_b200.l2norm(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
nbytes = 2 * x.numel() * 2
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6546.6
print(json.dumps({"kernel": "l2norm_rows_kernel", "shape": [B, T, H, D], "dtype": "bf16", "ms": ms, "algorithmic_bytes": nbytes,
                  "achieved_GBps": nbytes / ms / 1e6, "peak_GBps": peak, "frac": nbytes / ms / 1e6 / peak}))
