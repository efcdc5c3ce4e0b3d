"""CPU tests of the oracle (no GPU): self-consistency, golden vectors, known answers.

Model: the test pyramid SURVEY.md section 4/7 asks for, since the reference ships no numerical tests.
"""
import glob
import os
import warnings

import numpy as np
import pytest
import torch

from oracle.gdr_ref import (chunk_schedule, gdr_backward_ref, gdr_chunk_ref, gdr_recurrent_ref, gdr_recurrent_varlen_ref, make_inputs,
                            max_rel_err)

import golden_util
from oracle.gdr_ref import gdr_chunk_backward_ref, per_clip_errors, per_frame_max_rel, rms_rel_err

GOLDEN = golden_util.FP32
_load = golden_util.load_fp32


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_recurrent(path):
    t, _ = _load(path)
    o, sT = gdr_recurrent_ref(t["q"], t["k"], t["v"], t["g"], t["beta"], None, t["s0"])
    assert max_rel_err(o, t["o"]) < 1e-5
    assert max_rel_err(sT, t["sT"]) < 1e-5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_chunked_and_c_port(path, c_oracle):
    t, C = _load(path)
    o, sT = gdr_chunk_ref(t["q"], t["k"], t["v"], t["g"], t["beta"], None, t["s0"], frame_tokens=C)
    assert max_rel_err(o, t["o"]) < 2e-5 and max_rel_err(sT, t["sT"]) < 2e-5
    o, sT = c_oracle.gdr_recurrent_c(t["q"], t["k"], t["v"], t["g"], t["beta"], None, t["s0"])
    assert max_rel_err(o, t["o"]) < 1e-5 and max_rel_err(sT, t["sT"]) < 1e-5


def test_golden_present():
    assert len(GOLDEN) >= 4 and len(golden_util.BF16) >= 6


@pytest.mark.parametrize("path", golden_util.BF16, ids=golden_util.ids(golden_util.BF16))
def test_bf16_golden_pins_the_oracle(path, c_oracle):
    """fla-naive on bf16-ROUNDED q, k, v (the values the tcgen05 kernel consumes; 256-frame clip and 1024-token frames
    included) against the C port (all cases) and the torch oracle (short cases)."""
    q, k, v, g, beta, S0, rows, o_rows, sT_ref, C = golden_util.load_bf16(path)
    o, sT = c_oracle.gdr_recurrent_c(q, k, v, g, beta, None, S0)
    assert max_rel_err(o[:, rows], o_rows) < 2e-5 and max_rel_err(sT, sT_ref) < 2e-5
    assert rms_rel_err(o[:, rows], o_rows) < 1e-5
    if q.shape[1] <= 600:
        o, sT = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
        assert max_rel_err(o[:, rows], o_rows) < 1e-5 and max_rel_err(sT, sT_ref) < 1e-5
        o, sT = gdr_chunk_ref(q, k, v, g, beta, None, S0, frame_tokens=C)
        assert max_rel_err(o[:, rows], o_rows) < 2e-5 and max_rel_err(sT, sT_ref) < 2e-5


def test_error_metrics():
    b = torch.zeros(2, 6, 2, 4)
    b[0] = 100.0
    b[1] = 0.01
    a = b.clone()
    a[1, 3, 1, 2] += 0.005                                   # a 50 % error on a small-magnitude chain
    assert max_rel_err(a, b) < 1e-4                          # ... invisible to the whole-tensor metric
    mo, ro, _ = per_clip_errors(a, b)
    assert mo[1, 1] == pytest.approx(0.5) and mo[0].max() == 0 and ro[1, 1] > 0.05
    curve = per_frame_max_rel(a, b, 2)
    assert curve.shape == (3,) and curve[1] == pytest.approx(0.5) and curve[0] == 0 and curve[2] == 0
    assert rms_rel_err(b, b) == 0 and rms_rel_err(2 * b, b) == pytest.approx(1.0)


def test_fla_naive_cross_check():
    """Independent third-party restatement (not a reference pin); skipped where fla is absent."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        naive = pytest.importorskip("fla.ops.gated_delta_rule.naive")
    q, k, v, g, beta, S0 = make_inputs(2, 2 * 49, 2, 64, 96, seed=5, frame_tokens=49)
    o, sT = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    of, sf = naive.naive_recurrent_gated_delta_rule(q, k, v, beta, g, initial_state=S0.clone(),
                                                    output_final_state=True)
    oc, sc = naive.naive_chunk_gated_delta_rule(q, k, v, g, beta, chunk_size=49, initial_state=S0.clone(),
                                                output_final_state=True)
    assert max_rel_err(o, of) < 1e-5 and max_rel_err(sT, sf) < 1e-5
    assert max_rel_err(o, oc) < 1e-5 and max_rel_err(sT, sc) < 1e-5


@pytest.mark.parametrize("frame_tokens,max_rows", [(49, 64), (0, 64), (0, 16), (256, 64), (49, 49), (0, 1)])
def test_chunk_equals_recurrent(frame_tokens, max_rows):
    T = 512 if frame_tokens == 256 else 3 * 49
    q, k, v, g, beta, S0 = make_inputs(2, T, 2, 32, 48, seed=7, frame_tokens=frame_tokens, correlated=frame_tokens > 0)
    o1, s1 = gdr_recurrent_ref(q, k, v, g, beta, 0.37, S0)
    o2, s2 = gdr_chunk_ref(q, k, v, g, beta, 0.37, S0, frame_tokens=frame_tokens, max_rows=max_rows)
    assert max_rel_err(o2, o1) < 2e-5 and max_rel_err(s2, s1) < 2e-5


def test_chunk_schedule():
    assert chunk_schedule(98, 49) == [(0, 49), (49, 49)]
    assert chunk_schedule(130, 0) == [(0, 64), (64, 64), (128, 2)]
    assert chunk_schedule(2048, 1024)[:3] == [(0, 64), (64, 64), (128, 64)] and len(chunk_schedule(2048, 1024)) == 32
    assert chunk_schedule(0, 0) == []


def test_state_carry_split():
    """F frames in one call == two calls chained through final_state (row a5)."""
    q, k, v, g, beta, S0 = make_inputs(2, 6 * 49, 2, 64, 64, seed=9, frame_tokens=49)
    o, sT = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    cut = 2 * 49
    oa, sa = gdr_recurrent_ref(q[:, :cut], k[:, :cut], v[:, :cut], g[:, :cut], beta[:, :cut], None, S0)
    ob, sb = gdr_recurrent_ref(q[:, cut:], k[:, cut:], v[:, cut:], g[:, cut:], beta[:, cut:], None, sa)
    assert torch.equal(torch.cat([oa, ob], 1), o) and torch.equal(sb, sT)


def test_pad_tokens_are_noops():
    """k=0, beta=0, g=0 tokens leave the state untouched (what the kernel's 49->64 padding relies on)."""
    q, k, v, g, beta, S0 = make_inputs(1, 20, 1, 32, 16, seed=3)
    o, sT = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    pad = lambda x, val=0.0: torch.cat([x, torch.full((1, 5) + tuple(x.shape[2:]), val)], 1)
    o2, sT2 = gdr_recurrent_ref(pad(q), pad(k), pad(v, 3.0), pad(g), pad(beta), None, S0)
    assert torch.equal(o2[:, :20], o) and torch.equal(sT2, sT)


# ---- analytic known-answer tests ----

def test_kat_delta_rule_overwrites():
    """g=0, beta=1, |k|=1: after writing (k, v), S^T k == v exactly."""
    torch.manual_seed(0)
    K, V = 16, 8
    k = torch.nn.functional.normalize(torch.randn(1, 1, 1, K), dim=-1)
    v = torch.randn(1, 1, 1, V)
    S0 = torch.randn(1, 1, K, V)
    _, S = gdr_recurrent_ref(k, k, v, torch.zeros(1, 1, 1), torch.ones(1, 1, 1), 1.0, S0)
    assert torch.allclose(torch.einsum("kv,k->v", S[0, 0], k[0, 0, 0]), v[0, 0, 0], atol=1e-5)


def test_kat_beta_zero_is_pure_decay():
    q, k, v, g, beta, S0 = make_inputs(1, 7, 2, 16, 8, seed=2)
    _, S = gdr_recurrent_ref(q, k, v, g, torch.zeros_like(beta), None, S0)
    expect = S0 * g.sum(1).exp()[..., None, None]
    assert torch.allclose(S, expect, rtol=1e-5, atol=1e-7)


def test_kat_orthonormal_keys():
    """Orthonormal keys, g=0, beta=1, S0=0: S = sum_i k_i v_i^T and reading q=k_j returns scale*v_j."""
    K, V = 8, 4
    keys = torch.eye(K)[None, :, None, :]                       # [1,K,1,K]
    vals = torch.arange(K * V, dtype=torch.float32).reshape(1, K, 1, V)
    o, S = gdr_recurrent_ref(keys, keys, vals, torch.zeros(1, K, 1), torch.ones(1, K, 1), 0.5, None)
    assert torch.equal(S[0, 0], vals[0, :, 0])
    assert torch.equal(o, 0.5 * vals)                            # read-after-write, token causal


def test_kat_hand_computed():
    """K=2, V=2, T=3 worked by hand from the recurrence (alpha=0.5 on token 1)."""
    q = torch.tensor([[1., 0.], [0., 1.], [1., 1.]]).reshape(1, 3, 1, 2)
    k = torch.tensor([[1., 0.], [0., 1.], [1., 0.]]).reshape(1, 3, 1, 2)
    v = torch.tensor([[2., 4.], [6., 8.], [1., 1.]]).reshape(1, 3, 1, 2)
    g = torch.tensor([0., np.log(0.5), 0.]).reshape(1, 3, 1)
    beta = torch.tensor([1., 0.5, 1.]).reshape(1, 3, 1)
    o, S = gdr_recurrent_ref(q, k, v, g, beta, 1.0, None)
    # t0: S=[[2,4],[0,0]], o=[2,4]
    # t1: S*=.5 -> [[1,2],[0,0]]; r=.5*([6,8]-[0,0])=[3,4]; S=[[1,2],[3,4]]; o=S^T[0,1]=[3,4]
    # t2: r=[1,1]-[1,2]=[0,-1]; S=[[1,1],[3,4]]; o=S^T[1,1]=[4,5]
    assert torch.allclose(o[0, :, 0], torch.tensor([[2., 4.], [3., 4.], [4., 5.]]), atol=1e-6)
    assert torch.allclose(S[0, 0], torch.tensor([[1., 1.], [3., 4.]]), atol=1e-6)


def test_linearity_in_values():
    """The map (v, S0) -> (o, S_T) is linear: scaling both by 2 scales the outputs by exactly 2."""
    q, k, v, g, beta, S0 = make_inputs(1, 50, 1, 16, 8, seed=4)
    o, S = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    o2, S2 = gdr_recurrent_ref(q, k, 2 * v, g, beta, None, 2 * S0)
    assert torch.equal(o2, 2 * o) and torch.equal(S2, 2 * S)


def test_varlen_oracle_is_the_per_clip_oracle():
    """Packed clips: equal lengths reproduce the batched oracle exactly; an empty clip passes its state through."""
    B, T, H, K, V = 3, 40, 2, 16, 24
    q, k, v, g, beta, S0 = make_inputs(B, T, H, K, V, seed=7)
    o_b, s_b = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    pk = lambda t: t.reshape(1, B * T, *t.shape[2:])
    cu = [0, T, 2 * T, 3 * T]
    o_p, s_p = gdr_recurrent_varlen_ref(pk(q), pk(k), pk(v), pk(g), pk(beta), cu, None, S0)
    assert torch.equal(o_p.reshape(B, T, H, V), o_b) and torch.equal(s_p, s_b)
    cu = [0, 25, 25, 3 * T]                      # clip 1 is empty, clips 0 and 2 are ragged
    o_r, s_r = gdr_recurrent_varlen_ref(pk(q), pk(k), pk(v), pk(g), pk(beta), cu, None, S0)
    assert torch.equal(s_r[1], S0[1])
    o0, s0 = gdr_recurrent_ref(pk(q)[:, :25], pk(k)[:, :25], pk(v)[:, :25], pk(g)[:, :25], pk(beta)[:, :25], None, S0[:1])
    assert torch.equal(o_r[:, :25], o0) and torch.equal(s_r[0], s0[0])


def test_backward_oracle_against_finite_differences():
    """gdr_backward_ref (the checker a backward kernel will be held to): directional derivatives of
    <do, o> + <dsT, S_T> by central differences in float64."""
    B, T, H, K, V = 2, 9, 2, 4, 5
    q, k, v, g, beta, S0 = (t.double() for t in make_inputs(B, T, H, K, V, seed=31))
    gen = torch.Generator().manual_seed(32)
    do = torch.randn(B, T, H, V, generator=gen, dtype=torch.float64)
    dsT = torch.randn(B, H, K, V, generator=gen, dtype=torch.float64)
    grads = gdr_backward_ref(q, k, v, g, beta, do, dsT, None, S0)

    def loss(q, k, v, g, beta, S0):
        S = S0.clone()
        tot = 0.0
        for i in range(T):
            S = S * g[:, i].exp()[..., None, None]
            r = v[:, i] - torch.einsum("bhkv,bhk->bhv", S, k[:, i])
            S = S + k[:, i][..., :, None] * (beta[:, i][..., None] * r)[..., None, :]
            tot = tot + (K ** -0.5 * torch.einsum("bhkv,bhk->bhv", S, q[:, i]) * do[:, i]).sum()
        return tot + (S * dsT).sum()

    args = [q, k, v, g, beta, S0]
    for idx, gr in enumerate(grads):
        d = torch.randn(args[idx].shape, generator=gen, dtype=torch.float64)
        eps = 1e-6
        plus = [a + eps * d if j == idx else a for j, a in enumerate(args)]
        minus = [a - eps * d if j == idx else a for j, a in enumerate(args)]
        fd = (loss(*plus) - loss(*minus)) / (2 * eps)
        an = (gr * d).sum()
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)), (idx, float(fd), float(an))


@pytest.mark.parametrize("T,C", [(150, 64), (64, 64), (37, 16)])
def test_chunked_backward_restatement_equals_autograd(T, C):
    """The chunk-wise backward algebra of the CUDA kernel (reverse scan over chunks, dS carried) against reverse-mode
    differentiation of the token recurrence, float64: all six gradients, with a final-state cotangent."""
    B, H, K, V = 2, 2, 16, 24
    q, k, v, g, beta, S0 = make_inputs(B, T, H, K, V, seed=3)
    gen = torch.Generator().manual_seed(4)
    do = torch.randn(B, T, H, V, generator=gen)
    dsT = torch.randn(B, H, K, V, generator=gen)
    ref = gdr_backward_ref(q, k, v, g, beta, do, dsT, None, S0)
    out = gdr_chunk_backward_ref(q, k, v, g, beta, do, dsT, None, S0, C=C)
    for name, a, b in zip("dq dk dv dg dbeta dS0".split(), out, ref):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-10, name
    ref0 = gdr_backward_ref(q, k, v, g, beta, do, None, None, None)
    out0 = gdr_chunk_backward_ref(q, k, v, g, beta, do, None, None, None, C=C)
    for name, a, b in zip("dq dk dv dg dbeta dS0".split(), out0, ref0):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-10, name
