"""CPU simulation: what a `kind::tf32` (10-bit mantissa operands, fp32 accumulation) version of the chunked algorithm could reach
against the fp32 recurrence -- the question behind "fp32 I/O on the tensor cores".  Operands of every product are rounded to tf32;
selected operands can be kept exact (= a hi/lo split, two products).  Prints max-rel errors of readout and final state for a
correlated EchoNet-shaped case, a CAMUS-shaped case and BASELINE configs[0].  Result: 0.6e-3 .. 1.4e-3 with
plain tf32, 0.5e-3 .. 1.1e-3 with the state, V_new and the solve's operands split -- around the 1e-3 bound, not inside it, because
the partner operand of every product is still rounded; only a full 3xTF32 scheme (three products per product) would be."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.gdr_ref import gdr_recurrent_ref, make_inputs, max_rel_err, chunk_schedule


def tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


def mm(a, b, exact_a=False, exact_b=False):
    if exact_a:
        ah = tf32(a); return ah @ tf32(b) + tf32(a - ah) @ tf32(b)
    if exact_b:
        bh = tf32(b); return tf32(a) @ bh + tf32(a) @ tf32(b - bh)
    return tf32(a) @ tf32(b)


def chunked(q, k, v, g, beta, S0, C, cfg):
    B, T, H, K = k.shape
    V = v.shape[-1]
    scale = 1 / math.sqrt(K)
    f = lambda x: x.float().permute(0, 2, 1, 3) if x.dim() == 4 else x.float().permute(0, 2, 1)
    q, k, v = map(f, (q, k, v)); g, beta = f(g), f(beta)
    S = S0.clone().float()
    o = torch.empty(B, H, T, V)
    for (t0, n) in chunk_schedule(T, 0, C):
        sl = slice(t0, t0 + n)
        Q, Kc, Vc = q[:, :, sl], k[:, :, sl], v[:, :, sl]
        G = g[:, :, sl].cumsum(-1); bt = beta[:, :, sl]; e = G.exp()
        low = torch.tril(torch.ones(n, n, dtype=torch.bool))
        D = torch.where(low, G[..., :, None] - G[..., None, :], torch.full((n, n), -float("inf"))).exp()
        KK = Kc @ Kc.transpose(-1, -2) if cfg["solve"] else mm(Kc, Kc.transpose(-1, -2))
        A = torch.tril(bt[..., None] * KK * D, -1)
        eye = torch.eye(n).expand_as(A)
        Tm = torch.linalg.solve_triangular(eye + A, eye.clone(), upper=False)
        W = mm(Tm, (bt * e)[..., None] * Kc, exact_a=cfg["solve"]); U = mm(Tm, bt[..., None] * Vc, exact_a=cfg["solve"])
        Vn = U - mm(W, S, exact_b=cfg["state"])
        P = mm(Q, Kc.transpose(-1, -2)) * D
        o[:, :, sl] = scale * (mm(e[..., None] * Q, S, exact_b=cfg["state"]) + mm(P, Vn, exact_b=cfg["vnew"]))
        S = e[..., -1:, None] * S + mm((Kc * (e[..., -1:] / e)[..., None]).transpose(-1, -2), Vn, exact_b=cfg["vnew"])
    return o.permute(0, 2, 1, 3), S


if __name__ == "__main__":
    data = []
    for (B, T, H, V, C, corr) in ((2, 5 * 49, 3, 256, 49, True), (1, 2 * 1024, 2, 256, 1024, True), (1, 32 * 49, 1, 256, 49, False)):
        inp = make_inputs(B, T, H, 64, V, seed=100 + T, frame_tokens=C, correlated=corr)
        data.append((inp, gdr_recurrent_ref(*inp[:5], None, inp[5]), C))
    for cfg in (dict(solve=0, state=0, vnew=0), dict(solve=0, state=1, vnew=1), dict(solve=1, state=1, vnew=1)):
        res = []
        for inp, (o_ref, s_ref), C in data:
            o, s = chunked(*inp, C, cfg)
            res.append(f"readout {max_rel_err(o, o_ref):.1e} state {max_rel_err(s, s_ref):.1e}")
        print("split operands:", [k for k, v in cfg.items() if v] or "none", "|", " | ".join(res))
