"""Device-timed sweep over forced time-segment counts (flags bits 8-11) on one bench shape.
Usage (on a B200): python scripts/seg_sweep.py [B] [frames] [C] [H] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import gdkvm_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
C = int(sys.argv[3]) if len(sys.argv) > 3 else 49
H = int(sys.argv[4]) if len(sys.argv) > 4 else 8
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
T, K, V = F * C, 64, 256
g = torch.Generator(device="cuda").manual_seed(1)
rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
l2 = lambda x: torch.nn.functional.normalize(x, dim=-1)
q, k, v = l2(rn(B, T, H, K)).bfloat16(), l2(rn(B, T, H, K)).bfloat16(), rn(B, T, H, V).bfloat16()
beta, gate, S0 = torch.sigmoid(rn(B, T, H)), torch.nn.functional.logsigmoid(rn(B, T, H) + 4.0), 0.1 * rn(B, H, K, V)
o, sT = torch.empty_like(v), torch.empty_like(S0)
out = {}
for rnd in range(int(os.environ.get("ROUNDS", "2"))):
    for n in [int(x) for x in os.environ.get("SEGS", "1,0,2,3,4,5,6,7,8").split(",")]:
        fl = n << 8
        for _ in range(3):
            gdkvm_b200.gdr_lkva_out(q, k, v, gate, beta, o, sT, None, S0, C, fl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            gdkvm_b200.gdr_lkva_out(q, k, v, gate, beta, o, sT, None, S0, C, fl)
        e1.record()
        torch.cuda.synchronize()
        out.setdefault(n, []).append(round(e0.elapsed_time(e1) / reps, 4))
print(json.dumps({"shape": [B, F, C, H], "auto": gdkvm_b200.plan_segments(q, k, v, gate, beta, frame_tokens=C), "ms_by_segments(0=auto)": out}))
