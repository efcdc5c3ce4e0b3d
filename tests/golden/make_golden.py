"""Generates tests/golden/*.npz -- run HERE (CPU container), commit the outputs.

The reference tree has no golden vectors for this path, so the fixtures pin
the oracle against an INDEPENDENT pure-PyTorch implementation that happens to be installed in this
image: flash-linear-attention 0.5.1 ``naive_recurrent_gated_delta_rule``
(fla/ops/gated_delta_rule/naive.py:13).  That package is third-party, not a reference pin, and is
not available to the GPU box's tests by contract -- hence committed vectors.

    python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.gdr_ref import make_inputs  # noqa: E402

CASES = {
    # name: (B, T, H, K, V, frame_tokens, correlated, seed)
    "echonet_small": (1, 4 * 49, 2, 64, 64, 49, False, 11),
    "echonet_corr": (2, 3 * 49, 1, 64, 128, 49, True, 12),
    "ragged_tail": (1, 37, 1, 32, 40, 0, False, 13),
    "camus_small": (1, 2 * 256, 1, 64, 64, 256, True, 14),
}


# bf16-rounded q, k, v (what the tcgen05 chunk kernel consumes): fla-naive evaluated in fp32 on exactly those values.  The
# inputs are NOT stored -- `make_inputs(seed=...)` regenerates them (CPU torch generator; float64 checksums are stored and
# asserted by the test) -- and for the long clips only every `keep`-th frame of the readout is kept, so the fixtures stay
# small.  Key: name -> (B, T, H, K, V, frame_tokens, correlated, seed, keep-every-n-frames)
BF16_CASES = {
    "bf16_echonet_small": (1, 4 * 49, 2, 64, 64, 49, False, 11, 1),
    "bf16_echonet_corr": (2, 3 * 49, 1, 64, 128, 49, True, 12, 1),
    "bf16_camus_small": (1, 2 * 256, 1, 64, 64, 256, True, 14, 1),
    "bf16_echonet_v256": (1, 6 * 49, 2, 64, 256, 49, True, 15, 1),
    "bf16_long_clip_256f": (1, 256 * 49, 1, 64, 64, 49, False, 16, 8),     # configs[3]: 256-frame clip, 196 chunks
    "bf16_camus_4f_1024": (1, 4 * 1024, 1, 64, 128, 1024, True, 17, 1),    # configs[2]: 1024-token frames
}


def bf16_case_inputs(name):
    B, T, H, K, V, C, corr, seed, keep = BF16_CASES[name]
    return make_inputs(B, T, H, K, V, seed=seed, frame_tokens=C, correlated=corr, dtype=torch.bfloat16)


def kept_rows(T, C, keep):
    """token rows of the frames 0, keep, 2 keep, ... and of the last frame"""
    F = T // C
    frames = sorted(set(range(0, F, keep)) | {F - 1})
    return torch.cat([torch.arange(f * C, (f + 1) * C) for f in frames])


def main_bf16():
    from fla.ops.gated_delta_rule.naive import naive_recurrent_gated_delta_rule
    for name, (B, T, H, K, V, C, corr, seed, keep) in BF16_CASES.items():
        q, k, v, g, beta, S0 = bf16_case_inputs(name)
        o, sT = naive_recurrent_gated_delta_rule(q.float(), k.float(), v.float(), beta, g, initial_state=S0.clone(),
                                                 output_final_state=True)
        rows = kept_rows(T, C, keep)
        chk = np.array([float(x.double().sum()) for x in (q, k, v, g, beta, S0)], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), o_rows=o[:, rows].numpy(), rows=rows.numpy().astype(np.int32),
                            sT=sT.numpy(), checksums=chk, frame_tokens=np.int32(C))
        print(name, tuple(o.shape), "kept rows", len(rows))


def main():
    warnings.simplefilter("ignore")
    from fla.ops.gated_delta_rule.naive import naive_recurrent_gated_delta_rule
    if "--bf16-only" in sys.argv:
        return main_bf16()
    for name, (B, T, H, K, V, C, corr, seed) in CASES.items():
        q, k, v, g, beta, S0 = make_inputs(B, T, H, K, V, seed=seed, frame_tokens=C, correlated=corr)
        o, sT = naive_recurrent_gated_delta_rule(q, k, v, beta, g, initial_state=S0.clone(),
                                                 output_final_state=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            q=q.numpy(), k=k.numpy(), v=v.numpy(), g=g.numpy(), beta=beta.numpy(),
                            s0=S0.numpy(), o=o.numpy(), sT=sT.numpy(), frame_tokens=np.int32(C))
        print(name, tuple(o.shape))
    main_bf16()


if __name__ == "__main__":
    main()
