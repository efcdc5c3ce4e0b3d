"""CPU oracle for the GDKVM memory hot path (LKVA readout + Gated Delta Rule state update).

TEST INFRASTRUCTURE ONLY.  Nothing under ``gdkvm_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker (never as the thing shipped or measured as "ours").

PARITY UNPINNED.  The mounted reference (``/root/reference``) contains no model code
(``README.md:1`` points at a separate repo, ``README.md:36-38`` "Quick Start TBD"), no golden
vectors and no numerical tests (``website/e2e/smoke.spec.ts:4-80`` are browser smoke tests), and
the paper PDF is not mounted (``.MISSING_LARGE_BLOBS:1``).  This oracle is therefore written
directly from the equations in ``BASELINE.json`` ``north_star`` / ``BASELINE.md`` §2, the only
reference text naming the concept being ``README.md:20`` and
``website/src/content/homepage/en.json:20``.  It is cross-checked (tests only, skip-if-missing)
against flash-linear-attention 0.5.1 ``fla/ops/gated_delta_rule/naive.py:13`` (recurrent) and
``:67`` (chunked), a third-party package that is *not* a reference pin.

Semantics (per clip b, head h; state S in R^{K x V}, fp32):

    for each token i in frame-major raster order:
        S   <- exp(g_i) * S                       # alpha_i gate
        r    = v_i - S^T k_i                      # delta-rule residual
        S   <- S + k_i (beta_i * r)^T             # == alpha S (I - beta k k^T) + beta v k^T (transposed)
        o_i  = scale * S^T q_i                    # LKVA readout, read-after-write

Layout: q,k [B,T,H,K]; v,o [B,T,H,V]; g,beta [B,T,H]; initial/final state [B,H,K,V].
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def _prep(q, k, v, g, beta, scale, initial_state):
    B, T, H, K = k.shape
    V = v.shape[-1]
    assert q.shape == (B, T, H, K) and v.shape == (B, T, H, V)
    assert g.shape == (B, T, H) and beta.shape == (B, T, H)
    if scale is None:
        scale = 1.0 / math.sqrt(K)
    f = lambda x: x.detach().to("cpu", torch.float32)
    q, k, v, g, beta = map(f, (q, k, v, g, beta))
    if initial_state is None:
        S = torch.zeros(B, H, K, V, dtype=torch.float32)
    else:
        assert initial_state.shape == (B, H, K, V)
        S = f(initial_state).clone()
    return q, k, v, g, beta, float(scale), S, (B, T, H, K, V)


def gdr_recurrent_ref(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    g: torch.Tensor,
    beta: torch.Tensor,
    scale: Optional[float] = None,
    initial_state: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Token-recurrent ground truth, fp32 on CPU.

    Follows BASELINE.md §2 line by line; one python iteration per token, vectorised over
    (clip, head).  Returns ``(o [B,T,H,V] fp32, final_state [B,H,K,V] fp32)``.
    """
    q, k, v, g, beta, scale, S, (B, T, H, K, V) = _prep(q, k, v, g, beta, scale, initial_state)
    o = torch.empty(B, T, H, V, dtype=torch.float32)
    for i in range(T):
        k_i = k[:, i]                                   # [B,H,K]
        S = S * g[:, i].exp()[..., None, None]          # alpha gate
        r = v[:, i] - torch.einsum("bhkv,bhk->bhv", S, k_i)
        S = S + k_i[..., :, None] * (beta[:, i][..., None] * r)[..., None, :]
        o[:, i] = scale * torch.einsum("bhkv,bhk->bhv", S, q[:, i])
    return o, S


def chunk_schedule(T: int, frame_tokens: int = 0, max_rows: int = 64):
    """Chunk boundaries [(start, n_valid)] the kernel uses.

    ``frame_tokens == 0``: tile the flat token stream in ``max_rows`` tokens.
    ``frame_tokens == C``: every frame is one chunk, split in ``max_rows`` sub-chunks when
    ``C > max_rows`` (CAMUS: 1024 tokens -> 16 x 64); a chunk never straddles two frames.
    """
    out = []
    if T == 0:
        return out
    C = T if frame_tokens <= 0 else frame_tokens
    assert T % C == 0, "T must be a whole number of frames"
    for f0 in range(0, T, C):
        for c0 in range(0, C, max_rows):
            out.append((f0 + c0, min(max_rows, C - c0)))
    return out


def gdr_chunk_ref(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    g: torch.Tensor,
    beta: torch.Tensor,
    scale: Optional[float] = None,
    initial_state: Optional[torch.Tensor] = None,
    frame_tokens: int = 0,
    max_rows: int = 64,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Chunked WY/UT restatement, fp32 on CPU.

    Documents the math the sm_100a kernel executes; exactly equal to ``gdr_recurrent_ref`` in
    exact arithmetic for any chunk schedule.  Per chunk with Gamma = cumsum(g):
        A  = strict_tril(diag(beta) K K^T * exp(Gamma_i - Gamma_j));  T = (I + A)^-1
        W  = T diag(beta) (K * exp(Gamma));   U = T diag(beta) V;      V_new = U - W S
        O  = scale [ (Q * exp(Gamma)) S + tril(Q K^T * exp(Gamma_i - Gamma_j)) V_new ]
        S <- exp(Gamma_last) S + (K * exp(Gamma_last - Gamma))^T V_new
    """
    q, k, v, g, beta, scale, S, (B, T, H, K, V) = _prep(q, k, v, g, beta, scale, initial_state)
    o = torch.empty(B, T, H, V, dtype=torch.float32)
    # work in [B,H,t,*]
    qh, kh, vh = (x.permute(0, 2, 1, 3) for x in (q, k, v))
    gh, bh = g.permute(0, 2, 1), beta.permute(0, 2, 1)
    for (t0, n) in chunk_schedule(T, frame_tokens, max_rows):
        sl = slice(t0, t0 + n)
        Q, Kc, Vc = qh[:, :, sl], kh[:, :, sl], vh[:, :, sl]
        G = gh[:, :, sl].cumsum(-1)                       # [B,H,n]
        bt = bh[:, :, sl]
        diff = G[..., :, None] - G[..., None, :]          # Gamma_i - Gamma_j
        low = torch.tril(torch.ones(n, n, dtype=torch.bool))
        decay = torch.where(low, diff, torch.full_like(diff, -float("inf"))).exp()   # mask before exp
        KK = Kc @ Kc.transpose(-1, -2)
        A = torch.tril(bt[..., :, None] * KK * decay, diagonal=-1)
        eye = torch.eye(n, dtype=torch.float32).expand_as(A)
        Tm = torch.linalg.solve_triangular(eye + A, eye.clone(), upper=False)
        Kg = Kc * G.exp()[..., None]
        W = Tm @ (bt[..., None] * Kg)
        U = Tm @ (bt[..., None] * Vc)
        Vn = U - W @ S
        P = torch.tril((Q @ Kc.transpose(-1, -2)) * decay)
        Oc = scale * ((Q * G.exp()[..., None]) @ S + P @ Vn)
        o[:, sl] = Oc.permute(0, 2, 1, 3)
        Gl = G[..., -1:]
        S = Gl.exp()[..., None] * S + (Kc * (Gl - G).exp()[..., None]).transpose(-1, -2) @ Vn
    return o, S


def gdr_backward_ref(q, k, v, g, beta, do, dsT=None, scale=None, initial_state=None, dtype=torch.float64):
    """Gradients of the token recurrence by reverse-mode differentiation of the same four lines (ground truth for the
    backward pass, SURVEY.md section 8f rank 1: the checker ``csrc/gdr_bwd_sm100.cu`` is held to).

    ``do`` [B,T,H,V] and ``dsT`` [B,H,K,V] (optional) are the cotangents of the readout and of the final state.  Returns
    ``(dq, dk, dv, dg, dbeta, dS0)`` in ``dtype`` (float64 by default: the recurrence is re-run in that precision).
    """
    B, T, H, K = k.shape
    V = v.shape[-1]
    if scale is None:
        scale = 1.0 / math.sqrt(K)
    leaf = lambda x: x.detach().to("cpu", dtype).clone().requires_grad_(True)
    q_, k_, v_, g_, b_ = map(leaf, (q, k, v, g, beta))
    S0 = leaf(initial_state) if initial_state is not None else torch.zeros(B, H, K, V, dtype=dtype, requires_grad=True)
    S, outs = S0, []
    for i in range(T):
        k_i = k_[:, i]
        S = S * g_[:, i].exp()[..., None, None]
        r = v_[:, i] - torch.einsum("bhkv,bhk->bhv", S, k_i)
        S = S + k_i[..., :, None] * (b_[:, i][..., None] * r)[..., None, :]
        outs.append(float(scale) * torch.einsum("bhkv,bhk->bhv", S, q_[:, i]))
    o = torch.stack(outs, 1)
    loss = (o * do.detach().to("cpu", dtype)).sum()
    if dsT is not None:
        loss = loss + (S * dsT.detach().to("cpu", dtype)).sum()
    return torch.autograd.grad(loss, (q_, k_, v_, g_, b_, S0))


def gdr_chunk_backward_ref(q, k, v, g, beta, do, dsT=None, scale=None, S0=None, C=64, dt=torch.float64):
    """Chunked (WY/UT) restatement of the backward pass -- the algebra ``csrc/gdr_bwd_sm100.cu`` executes, chunk by chunk in
    reverse time with the state cotangent ``dS`` carried across chunks; checked against ``gdr_backward_ref`` (autograd
    through the token recurrence) in tests/test_oracle.py.  Flat ``C``-token chunks.  Returns (dq, dk, dv, dg, dbeta, dS0)."""
    B, T, H, K = k.shape; V = v.shape[-1]
    scale = scale or 1 / math.sqrt(K)
    f = lambda x: x.to(dt).permute(0, 2, 1, 3) if x.dim() == 4 else x.to(dt).permute(0, 2, 1)
    q, k, v, do = map(f, (q, k, v, do)); g, beta = f(g), f(beta)
    S = torch.zeros(B, H, K, V, dtype=dt) if S0 is None else S0.to(dt).clone()
    sched = chunk_schedule(T, 0, C)
    # forward: store chunk-start states and Vn
    Ss, Vns = [], []
    for (t0, n) in sched:
        sl = slice(t0, t0+n)
        Q, Kc, Vc = q[:, :, sl], k[:, :, sl], v[:, :, sl]
        G = g[:, :, sl].cumsum(-1); bt = beta[:, :, sl]; e = G.exp()
        low = torch.tril(torch.ones(n, n, dtype=torch.bool))
        D = torch.where(low, G[..., :, None] - G[..., None, :], torch.full((n, n), -float('inf'), dtype=dt)).exp()
        A = torch.tril(bt[..., None] * (Kc @ Kc.transpose(-1, -2)) * D, -1)
        eye = torch.eye(n, dtype=dt).expand_as(A)
        Tm = torch.linalg.solve_triangular(eye + A, eye.clone(), upper=False)
        W = Tm @ (bt[..., None] * e[..., None] * Kc); U = Tm @ (bt[..., None] * Vc)
        Vn = U - W @ S
        Ss.append(S); Vns.append(Vn)
        S = e[..., -1:, None] * S + (Kc * (e[..., -1:] / e)[..., None]).transpose(-1, -2) @ Vn
    dS = torch.zeros_like(S) if dsT is None else dsT.to(dt).clone()
    dq = torch.zeros_like(q); dk = torch.zeros_like(k); dv = torch.zeros_like(v); dg = torch.zeros_like(g); db = torch.zeros_like(beta)
    for ci in range(len(sched) - 1, -1, -1):
        t0, n = sched[ci]; sl = slice(t0, t0+n)
        Q, Kc, Vc, dO = q[:, :, sl], k[:, :, sl], v[:, :, sl], do[:, :, sl]
        S, Vn = Ss[ci], Vns[ci]
        G = g[:, :, sl].cumsum(-1); bt = beta[:, :, sl]; e = G.exp(); gam = e[..., -1]
        low = torch.tril(torch.ones(n, n, dtype=torch.bool))
        D = torch.where(low, G[..., :, None] - G[..., None, :], torch.full((n, n), -float('inf'), dtype=dt)).exp()
        KK = Kc @ Kc.transpose(-1, -2)
        A = torch.tril(bt[..., None] * KK * D, -1)
        eye = torch.eye(n, dtype=dt).expand_as(A)
        Tm = torch.linalg.solve_triangular(eye + A, eye.clone(), upper=False)
        Bt = (bt * e)[..., None] * Kc; Vt = bt[..., None] * Vc
        W = Tm @ Bt
        P = (Q @ Kc.transpose(-1, -2)) * D           # tril via D
        Kh = (gam[..., None] / e)[..., None] * Kc
        dVn = scale * P.transpose(-1, -2) @ dO + Kh @ dS
        dSn = gam[..., None, None] * dS + scale * (e[..., None] * Q).transpose(-1, -2) @ dO - W.transpose(-1, -2) @ dVn
        dP = scale * (dO @ Vn.transpose(-1, -2)) * low
        dQS = scale * dO @ S.transpose(-1, -2)        # C x K
        dPD = dP * D
        dQ = e[..., None] * dQS + dPD @ Kc
        dK = dPD.transpose(-1, -2) @ Q
        dKh = Vn @ dS.transpose(-1, -2)
        dK = dK + (gam[..., None] / e)[..., None] * dKh
        dW = -dVn @ S.transpose(-1, -2)
        dT = dW @ Bt.transpose(-1, -2) + dVn @ Vt.transpose(-1, -2)
        dBt = Tm.transpose(-1, -2) @ dW; dVt = Tm.transpose(-1, -2) @ dVn
        dV = bt[..., None] * dVt
        dK = dK + (bt * e)[..., None] * dBt
        dbeta = e * (dBt * Kc).sum(-1) + (dVt * Vc).sum(-1)
        dA = -(Tm.transpose(-1, -2) @ dT @ Tm.transpose(-1, -2)) * torch.tril(torch.ones(n, n, dtype=dt), -1)
        dbeta = dbeta + (dA * KK * D).sum(-1)
        M = dA * bt[..., None] * D
        dK = dK + M @ Kc + M.transpose(-1, -2) @ Kc
        dGam = e * (Q * dQS).sum(-1) + (dP * P).sum(-1) - (dP * P).sum(-2) + (dA * A).sum(-1) - (dA * A).sum(-2) \
               + (dBt * Bt).sum(-1) - (dKh * Kh).sum(-1)
        last = (dKh * Kh).sum((-1, -2)) + gam * (dS * S).sum((-1, -2))
        dGam[..., -1] += last
        dg[:, :, sl] = dGam.flip(-1).cumsum(-1).flip(-1)
        dq[:, :, sl] = dQ; dk[:, :, sl] = dK; dv[:, :, sl] = dV; db[:, :, sl] = dbeta
        dS = dSn
    p = lambda x: x.permute(0, 2, 1, 3) if x.dim() == 4 else x.permute(0, 2, 1)
    return p(dq), p(dk), p(dv), p(dg), p(db), dS


def gdr_recurrent_varlen_ref(q, k, v, g, beta, cu_seqlens, scale=None, initial_state=None):
    """Packed variable-length clips (fla's ``cu_seqlens``, fla/ops/gated_delta_rule/chunk.py:375): q,k,v [1,T,H,*], clip n =
    rows cu_seqlens[n] .. cu_seqlens[n+1]-1, states [N,H,K,V].  One ``gdr_recurrent_ref`` call per clip; a clip without
    tokens passes its initial state through.  Returns (o [1,T,H,V] fp32, final_state [N,H,K,V] fp32)."""
    N = len(cu_seqlens) - 1
    H, K, V = q.shape[2], q.shape[3], v.shape[3]
    o = torch.zeros(1, q.shape[1], H, V, dtype=torch.float32)
    sT = torch.zeros(N, H, K, V, dtype=torch.float32)
    for n in range(N):
        a, b = int(cu_seqlens[n]), int(cu_seqlens[n + 1])
        s0 = initial_state[n:n + 1] if initial_state is not None else None
        if b == a:
            if s0 is not None:
                sT[n] = s0[0]
            continue
        o_n, s_n = gdr_recurrent_ref(q[:, a:b], k[:, a:b], v[:, a:b], g[:, a:b], beta[:, a:b], scale, s0)
        o[:, a:b] = o_n
        sT[n] = s_n[0]
    return o, sT


def make_inputs(B, T, H, K, V, *, seed=1234, frame_tokens=0, correlated=False,
                dtype=torch.float32, with_state=True):
    """Seeded synthetic inputs of SURVEY.md §8(d) / BASELINE.md §3.

    q,k = l2norm(randn); v = randn; beta = sigmoid(randn); g = logsigmoid(randn + 4);
    S0 = 0.1 randn.  ``correlated=True`` draws one base key per frame and adds 0.3 randn
    (adjacent pixels of one frame are highly correlated -- the hard case for the solve).
    q,k,v are rounded to ``dtype`` (bf16 configs) so oracle and kernel see the same values.
    """
    gen = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float32)
    l2 = lambda x: x / x.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    q = l2(rn(B, T, H, K))
    if correlated and frame_tokens > 0:
        F = T // frame_tokens
        base = rn(B, F, 1, H, K).expand(B, F, frame_tokens, H, K).reshape(B, T, H, K)
        k = l2(base + 0.3 * rn(B, T, H, K))
    else:
        k = l2(rn(B, T, H, K))
    v = rn(B, T, H, V)
    beta = torch.sigmoid(rn(B, T, H))
    g = torch.nn.functional.logsigmoid(rn(B, T, H) + 4.0)
    S0 = 0.1 * rn(B, H, K, V) if with_state else None
    q, k, v = (x.to(dtype) for x in (q, k, v))
    return q, k, v, g, beta, S0


def max_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Tolerance metric of BASELINE.md §2: max|a-b| / max|b| per tensor."""
    a = a.detach().to("cpu", torch.float32)
    b = b.detach().to("cpu", torch.float32)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rms_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Root-mean-square error relative to the reference's RMS: ||a-b||_2 / ||b||_2 per tensor.  Unlike ``max_rel_err``
    it is not set by one outlier element, and unlike a per-element relative error it is defined where b crosses zero."""
    a = a.detach().to("cpu", torch.float64)
    b = b.detach().to("cpu", torch.float64)
    return float((a - b).square().mean().sqrt() / b.square().mean().sqrt().clamp_min(1e-30))


def per_clip_errors(o, o_ref, sT=None, s_ref=None):
    """[(clip, head) -> max-rel of the readout, rms-rel of the readout, max-rel of the final state] for every chain of a
    batched call (o [B,T,H,V], states [B,H,K,V]): a chain with small-magnitude outputs must not hide behind the batch's
    largest element, which is what a whole-tensor ``max_rel_err`` allows."""
    o = o.detach().to("cpu", torch.float32)
    o_ref = o_ref.detach().to("cpu", torch.float32)
    B, T, H, V = o_ref.shape
    d = (o - o_ref)
    mo = d.abs().amax(dim=(1, 3)) / o_ref.abs().amax(dim=(1, 3)).clamp_min(1e-30)                         # [B,H]
    ro = d.double().square().mean(dim=(1, 3)).sqrt() / o_ref.double().square().mean(dim=(1, 3)).sqrt().clamp_min(1e-30)
    ms = None
    if sT is not None:
        sT = sT.detach().to("cpu", torch.float32)
        s_ref = s_ref.detach().to("cpu", torch.float32)
        ms = (sT - s_ref).abs().amax(dim=(2, 3)) / s_ref.abs().amax(dim=(2, 3)).clamp_min(1e-30)
    return mo, ro.float(), ms


def per_frame_max_rel(o, o_ref, frame_tokens: int):
    """Error growth along a clip: for every frame f, max over (clip, head) of max|a-b| / max|b| taken over that frame's
    rows of that chain only.  Returns a 1-D tensor of length T / frame_tokens."""
    o = o.detach().to("cpu", torch.float32)
    o_ref = o_ref.detach().to("cpu", torch.float32)
    B, T, H, V = o_ref.shape
    F = T // frame_tokens
    a = o[:, :F * frame_tokens].reshape(B, F, frame_tokens, H, V)
    b = o_ref[:, :F * frame_tokens].reshape(B, F, frame_tokens, H, V)
    e = (a - b).abs().amax(dim=(2, 4)) / b.abs().amax(dim=(2, 4)).clamp_min(1e-30)                         # [B,F,H]
    return e.amax(dim=(0, 2))
