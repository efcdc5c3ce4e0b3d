"""One small chunk-kernel launch for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from oracle.gdr_ref import gdr_recurrent_ref, make_inputs, max_rel_err
for (B, T, H, V, C) in ((2, 4 * 49, 2, 256, 49), (1, 130, 1, 128, 0)):
    q, k, v, g, beta, S0 = make_inputs(B, T, H, 64, V, seed=5, frame_tokens=C, dtype=torch.bfloat16)
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    o, sT = gdkvm_b200.gdr_lkva(q.cuda(), k.cuda(), v.cuda(), g.cuda(), beta.cuda(), None, S0.cuda(), True, C, 2)
    torch.cuda.synchronize()
    print("case", (B, T, H, V, C), "o", max_rel_err(o, o_ref), "S", max_rel_err(sT, s_ref), flush=True)
