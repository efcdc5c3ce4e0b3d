"""Loading of tests/golden/*.npz (see tests/golden/make_golden.py, which wrote them in this container with
flash-linear-attention's naive_recurrent_gated_delta_rule; the GPU box has neither fla nor /root/reference)."""
import glob
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL = sorted(glob.glob(os.path.join(HERE, "*.npz")))
FP32 = [p for p in ALL if not os.path.basename(p).startswith("bf16_")]     # inputs stored, fp32 q/k/v
BF16 = [p for p in ALL if os.path.basename(p).startswith("bf16_")]         # inputs regenerated from the seed, bf16 q/k/v
ids = lambda paths: [os.path.basename(p)[:-4] for p in paths]


def _maker():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_fp32(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in ("q", "k", "v", "g", "beta", "s0", "o", "sT")}, int(z["frame_tokens"])


def load_bf16(path):
    """-> (q, k, v (bf16), g, beta, S0, rows, o_rows, sT, frame_tokens): the readout of the golden is given on `rows` only.
    The inputs come from the seeded generator; their float64 checksums must equal the ones stored with the outputs."""
    z = np.load(path)
    name = os.path.basename(path)[:-4]
    q, k, v, g, beta, S0 = _maker().bf16_case_inputs(name)
    chk = np.array([float(x.double().sum()) for x in (q, k, v, g, beta, S0)], dtype=np.float64)
    assert np.allclose(chk, z["checksums"], rtol=1e-12, atol=1e-9), f"{name}: regenerated inputs differ from the generator run"
    return (q, k, v, g, beta, S0, torch.from_numpy(z["rows"]).long(), torch.from_numpy(z["o_rows"]), torch.from_numpy(z["sT"]),
            int(z["frame_tokens"]))
