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


def main():
    warnings.simplefilter("ignore")
    from fla.ops.gated_delta_rule.naive import naive_recurrent_gated_delta_rule
    for name, (B, T, H, K, V, C, corr, seed) in CASES.items():
        q, k, v, g, beta, S0 = make_inputs(B, T, H, K, V, seed=seed, frame_tokens=C, correlated=corr)
        o, sT = naive_recurrent_gated_delta_rule(q, k, v, beta, g, initial_state=S0.clone(),
                                                 output_final_state=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            q=q.numpy(), k=k.numpy(), v=v.numpy(), g=g.numpy(), beta=beta.numpy(),
                            s0=S0.numpy(), o=o.numpy(), sT=sT.numpy(), frame_tokens=np.int32(C))
        print(name, tuple(o.shape))


if __name__ == "__main__":
    main()
