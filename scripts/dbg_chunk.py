"""Quick parity read-out of the chunked kernel on a few shapes (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdkvm_b200
from oracle.gdr_ref import gdr_recurrent_ref, make_inputs, max_rel_err

cases = [  # B,T,H,K,V,C,corr
    (1, 64, 1, 64, 128, 0, False),
    (1, 64, 1, 64, 256, 0, False),
    (1, 128, 1, 64, 256, 0, False),
    (2, 4 * 49, 2, 64, 256, 49, False),
    (2, 4 * 49, 2, 64, 256, 49, True),
    (1, 2 * 256, 1, 64, 256, 256, True),
    (1, 130, 2, 64, 128, 0, False),
    (1, 32 * 49, 1, 64, 256, 49, False),
]
import itertools
for (B, T, H, K, V, C, corr), gmul in itertools.chain(zip(cases, itertools.repeat(1.0)), [((2, 4 * 49, 2, 64, 256, 49, True), 100.0), ((1, 256, 1, 64, 128, 0, False), 300.0)]):
    q, k, v, g, beta, S0 = make_inputs(B, T, H, K, V, seed=T + V, frame_tokens=C, correlated=corr, dtype=torch.bfloat16)
    g = g * gmul
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    for name, flags in (("frame", 2), ("flat", 6)):
        try:
            o, sT = gdkvm_b200.gdr_lkva(q.cuda(), k.cuda(), v.cuda(), g.cuda(), beta.cuda(), None, S0.cuda(), True, C, flags)
            torch.cuda.synchronize()
            print(f"B{B} T{T} H{H} V{V} C{C} corr{int(corr)} gmul{gmul:g} {name}: o {max_rel_err(o, o_ref):.3e}  S {max_rel_err(sT, s_ref):.3e}", flush=True)
        except Exception as e:
            print(f"B{B} T{T} H{H} V{V} C{C} {name}: EXC {e}", flush=True)
            raise
