"""CPU model of the tcgen05 chunk kernel's ARITHMETIC (gdkvm_b200/csrc/gdr_chunked_sm100.cu):
same operand roundings (bf16 tensor-core inputs, fp32 accumulate, fp32 solve) in plain torch.
Test infrastructure: predicts the kernel's error against the oracle without a GPU and tells a
layout bug (error >> model) from a precision limit (error ~ model)."""
import math

import torch

ALL = ("Tp", "Kt", "Kp", "Qt", "P", "W", "Sb", "Vnb", "O")


def gdr_chunk_model(q, k, v, g, beta, scale=None, S0=None, frame_tokens=0, skip=(), split=()):
    """`skip`: operand names NOT rounded to bf16 (error attribution).  `split`: operands carried as
    a bf16 hi+lo pair (two MMAs), i.e. ~16 mantissa bits."""
    from oracle.gdr_ref import chunk_schedule
    B, T, H, K = k.shape
    V = v.shape[-1]
    scale = 1.0 / math.sqrt(K) if scale is None else scale

    def rd(name, x):
        if name in skip:
            return x
        hi = x.bfloat16().float()
        if name in split:
            return hi + (x - hi).bfloat16().float()
        return hi

    f32 = lambda x: x.float().permute(0, 2, 1, 3) if x.dim() == 4 else x.float().permute(0, 2, 1)
    qh, kh, vh, gh, bh = map(f32, (q, k, v, g, beta))
    S = torch.zeros(B, H, K, V) if S0 is None else S0.float().clone()
    o = torch.zeros(B, H, T, V)
    for (t0, n) in chunk_schedule(T, frame_tokens, 64):
        pad = lambda x: torch.nn.functional.pad(x[:, :, t0:t0 + n], (0, 0, 0, 64 - n))
        Q, Kc, Vc = pad(qh), pad(kh), pad(vh)
        gg = torch.nn.functional.pad(gh[:, :, t0:t0 + n], (0, 64 - n))
        bt = torch.nn.functional.pad(bh[:, :, t0:t0 + n], (0, 64 - n))
        G = gg.cumsum(-1)
        e = G.exp()
        dec = (G[..., :, None] - G[..., None, :]).clamp(max=0).exp()
        KK = Kc @ Kc.transpose(-1, -2)
        QK = Q @ Kc.transpose(-1, -2)
        A = torch.tril(bt[..., :, None] * KK * dec, -1)
        eye = torch.eye(64).expand_as(A)
        Tm = torch.linalg.solve_triangular(eye + A, eye.clone(), upper=False)
        Tp = rd("Tp", Tm * bt[..., None, :])
        Kt = rd("Kt", Kc * e[..., None])
        Kp = rd("Kp", Kc * (G[..., -1:] - G).exp()[..., None])
        Qt = rd("Qt", Q * (scale * e)[..., None])
        P = rd("P", torch.tril(scale * QK * dec))
        W = rd("W", Tp @ Kt)
        Sb = rd("Sb", S)
        Vn = Tp @ Vc - W @ Sb
        Vnb = rd("Vnb", Vn)
        Oc = Qt @ Sb + P @ Vnb
        o[:, :, t0:t0 + n] = rd("O", Oc)[:, :, :n]
        S = G[..., -1:].exp()[..., None] * S + Kp.transpose(-1, -2) @ Vnb
    return o.permute(0, 2, 1, 3), S
