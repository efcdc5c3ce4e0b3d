"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the torch.library op
and the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json north_star's: max|a-b|/max|b| <= 1e-3 with fp32 I/O (fp32 state),
<= 2e-2 with bf16 I/O, on the readout AND the final state.
"""
import glob
import os

import numpy as np
import pytest
import torch

import golden_util
from oracle.gdr_ref import (gdr_recurrent_ref, gdr_recurrent_varlen_ref, make_inputs, max_rel_err, per_clip_errors, per_frame_max_rel,
                            rms_rel_err)

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-3, torch.bfloat16: 2e-2}
RECURRENT, CHUNKED, FLAT, FRAME = 0x1, 0x2, 0x4, 0x8
GOLDEN = golden_util.FP32


def SEG(n):
    return (n & 0xF) << 8


@pytest.fixture(scope="module")
def op(built_lib):
    assert torch.cuda.is_available(), "-m gpu tests need a B200"
    import gdkvm_b200
    return gdkvm_b200


def _dev(*ts):
    return [t.cuda() if t is not None else None for t in ts]


def _run(op, q, k, v, g, beta, S0, **kw):
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    o, sT = op.gdr_lkva(qd, kd, vd, gd, bd, kw.pop("scale", None), sd, True, **kw)
    torch.cuda.synchronize()
    return o.float().cpu(), sT.cpu()


def _assert_per_chain(o, o_ref, sT, s_ref, tol, what=""):
    """The tolerance per (clip, head) chain -- max-rel of the readout and of the final state over that chain's own
    elements, so a chain with small outputs cannot hide behind the batch maximum -- and the RMS error well inside it."""
    mo, ro, ms = per_clip_errors(o, o_ref, sT, s_ref)
    assert float(mo.max()) <= tol, (what, "readout max-rel per chain", mo)
    assert ms is None or float(ms.max()) <= tol, (what, "final state max-rel per chain", ms)
    assert float(ro.max()) <= tol / 2, (what, "readout rms-rel per chain", ro)


def _paths(op, q, k, v, g, beta, C):
    """Every kernel path that can run this problem: forced recurrent, auto, and (if eligible) chunked."""
    paths = [("recurrent", dict(flags=RECURRENT)), ("auto", dict(flags=0))]
    if op.plan(q, k, v, g, beta, frame_tokens=C) == 1:
        paths += [("chunked", dict(flags=CHUNKED)), ("chunked_flat", dict(flags=CHUNKED | FLAT)),
                  ("chunked_frame", dict(flags=CHUNKED | FRAME))]
    return paths


@pytest.mark.parametrize("shape", [
    # B, T, H, K, V, frame_tokens, correlated
    (1, 32 * 49, 1, 64, 256, 49, False),     # BASELINE configs[0]
    (2, 5 * 49, 3, 64, 256, 49, True),       # correlated keys inside a frame
    (1, 2 * 1024, 2, 64, 256, 1024, True),   # CAMUS-shaped: 1024-token frames -> 16 sub-chunks
    (2, 130, 2, 64, 128, 0, False),          # ragged tail (130 = 2*64 + 2)
    (1, 37, 1, 32, 40, 0, False),            # K=32, V not a multiple of anything
    (1, 70, 2, 128, 128, 0, False),          # K=128
    (3, 1, 2, 64, 64, 0, False),             # single token
    (2, 5 * 49 + 13, 3, 64, 64, 0, True),    # V = 64: the chunk kernel with the upper TMEM lanes idle
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_parity_vs_oracle(op, shape, dtype):
    B, T, H, K, V, C, corr = shape
    q, k, v, g, beta, S0 = make_inputs(B, T, H, K, V, seed=100 + T, frame_tokens=C, correlated=corr, dtype=dtype)
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    for name, kw in _paths(op, q, k, v, g, beta, C):
        o, sT = _run(op, q, k, v, g, beta, S0, frame_tokens=C, **kw)
        eo, es = max_rel_err(o, o_ref), max_rel_err(sT, s_ref)
        assert eo <= TOL[dtype] and es <= TOL[dtype], (name, eo, es)
        _assert_per_chain(o, o_ref, sT, s_ref, TOL[dtype], name)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures(op, path):
    z = np.load(path)
    t = {n: torch.from_numpy(z[n]) for n in ("q", "k", "v", "g", "beta", "s0", "o", "sT")}
    C = int(z["frame_tokens"])
    for name, kw in _paths(op, t["q"], t["k"], t["v"], t["g"], t["beta"], C):
        o, sT = _run(op, t["q"], t["k"], t["v"], t["g"], t["beta"], t["s0"], frame_tokens=C, **kw)
        assert max_rel_err(o, t["o"]) <= 1e-3 and max_rel_err(sT, t["sT"]) <= 1e-3, name


@pytest.mark.parametrize("path", golden_util.BF16, ids=golden_util.ids(golden_util.BF16))
def test_bf16_golden_through_the_chunk_kernel(op, path):
    """The tcgen05 kernel, FORCED, against the independent goldens (fla-naive on the same bf16-rounded q, k, v; a 256-frame
    clip = 196 chunks and 1024-token frames included): frame-aligned, flat and time-segmented tiling, per chain, RMS too."""
    q, k, v, g, beta, S0, rows, o_rows, sT_ref, C = golden_util.load_bf16(path)
    assert op.plan(q.cuda(), k.cuda(), v.cuda(), g.cuda(), beta.cuda(), frame_tokens=C, flags=CHUNKED) == 1
    for fl in (CHUNKED, CHUNKED | FLAT, CHUNKED | FRAME, CHUNKED | SEG(3)):
        o, sT = _run(op, q, k, v, g, beta, S0, frame_tokens=C, flags=fl)
        eo, es = max_rel_err(o[:, rows], o_rows), max_rel_err(sT, sT_ref)
        assert eo <= 2e-2 and es <= 2e-2, (fl, eo, es)
        _assert_per_chain(o[:, rows], o_rows, sT, sT_ref, 2e-2, fl)
        assert rms_rel_err(o[:, rows], o_rows) <= 1e-2 and rms_rel_err(sT, sT_ref) <= 1e-2


def test_recurrent_kernel_is_fp32_exact(op):
    """The fp32 path follows the oracle operation for operation: far inside the 1e-3 contract."""
    q, k, v, g, beta, S0 = make_inputs(2, 4 * 49, 2, 64, 256, seed=3, frame_tokens=49)
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    o, sT = _run(op, q, k, v, g, beta, S0, flags=RECURRENT)
    assert max_rel_err(o, o_ref) < 2e-5 and max_rel_err(sT, s_ref) < 2e-5


def test_zero_state_and_no_final_state(op):
    q, k, v, g, beta, _ = make_inputs(1, 98, 2, 64, 128, seed=4, dtype=torch.bfloat16)
    o_ref, _ = gdr_recurrent_ref(q, k, v, g, beta, 0.2, None)
    qd, kd, vd, gd, bd = _dev(q, k, v, g, beta)
    o, sT = op.gdr_lkva(qd, kd, vd, gd, bd, 0.2, None, False)
    assert sT is None and max_rel_err(o, o_ref) <= 2e-2
    o2, sT2 = op.chunk_gated_delta_rule(qd, kd, vd, gd, bd, scale=0.2, output_final_state=True)
    assert torch.equal(o2, o) and sT2.shape == (1, 2, 64, 128)


def test_empty_clip(op):
    q, k, v, g, beta, S0 = make_inputs(2, 0, 2, 64, 64, seed=5)
    o, sT = _run(op, q, k, v, g, beta, S0)
    assert o.shape == (2, 0, 2, 64) and torch.equal(sT, S0)


def test_bf16_gates(op):
    q, k, v, g, beta, S0 = make_inputs(1, 98, 2, 64, 128, seed=6, dtype=torch.bfloat16)
    g, beta = g.bfloat16(), beta.bfloat16()
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    o, sT = _run(op, q, k, v, g, beta, S0)
    assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_state_carry_and_strided_views(op, dtype):
    """Two time segments (non-contiguous views of one clip batch) chained through final_state
    reproduce the single call (row a5, BASELINE configs[3] streaming)."""
    C = 49
    q, k, v, g, beta, S0 = make_inputs(3, 6 * C, 2, 64, 256, seed=8, frame_tokens=C, dtype=dtype)
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    # frame-aligned chunking (and the token-recurrent kernel) cuts at the same places in the
    # segmented and the single call; the default flat 64-token tiling does not (98 is not a multiple
    # of 64), so it is compared within the tolerance below instead of bit for bit
    for flags in (RECURRENT, FRAME):
        o, sT = op.gdr_lkva(qd, kd, vd, gd, bd, None, sd, True, C, flags)
        cut = 2 * C
        oa, sa = op.gdr_lkva(qd[:, :cut], kd[:, :cut], vd[:, :cut], gd[:, :cut], bd[:, :cut], None, sd, True, C, flags)
        ob, sb = op.gdr_lkva(qd[:, cut:], kd[:, cut:], vd[:, cut:], gd[:, cut:], bd[:, cut:], None, sa, True, C, flags)
        assert torch.equal(torch.cat([oa, ob], 1), o), flags     # identical chunking => bit-identical
        assert torch.equal(sb, sT), flags
    mem = op.GDRMemory(frame_tokens=C, flags=FRAME)
    o2, s2 = mem.forward_segments(qd, kd, vd, gd, bd, frames_per_segment=4, initial_state=sd)
    o1, s1 = mem(qd, kd, vd, gd, bd, sd)
    assert torch.equal(o2, o1) and torch.equal(s2, s1)
    mem = op.GDRMemory(frame_tokens=C)                 # default tiling: equal within the tolerance
    o3, s3 = mem.forward_segments(qd, kd, vd, gd, bd, frames_per_segment=4, initial_state=sd)
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    assert max_rel_err(o3, o_ref) <= TOL[dtype] and max_rel_err(s3, s_ref) <= TOL[dtype]


def test_strided_batch_views_on_the_chunk_path(op):
    """Every other clip of a larger batch (batch stride doubled, nothing copied) through the tcgen05 kernel."""
    q, k, v, g, beta, S0 = make_inputs(6, 3 * 64 + 9, 2, 64, 256, seed=31, dtype=torch.bfloat16)
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    sl = slice(None, None, 2)
    assert op.plan(qd[sl], kd[sl], vd[sl], gd[sl], bd[sl]) == 1
    o, sT = op.gdr_lkva(qd[sl], kd[sl], vd[sl], gd[sl], bd[sl], None, sd[sl].contiguous(), True, 0, CHUNKED)
    o_c, s_c = op.gdr_lkva(qd[sl].contiguous(), kd[sl].contiguous(), vd[sl].contiguous(), gd[sl].contiguous(),
                           bd[sl].contiguous(), None, sd[sl].contiguous(), True, 0, CHUNKED)
    assert torch.equal(o, o_c) and torch.equal(sT, s_c)
    o_ref, s_ref = gdr_recurrent_ref(q[sl], k[sl], v[sl], g[sl], beta[sl], None, S0[sl])
    assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2


def test_frame_aligned_128_token_frames_v128(op):
    """frame_tokens a multiple of 64 -> frame-aligned sub-chunks by default; V = 128 runs one state warpgroup."""
    q, k, v, g, beta, S0 = make_inputs(2, 3 * 128, 2, 64, 128, seed=32, frame_tokens=128, correlated=True, dtype=torch.bfloat16)
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    for flags in (CHUNKED, CHUNKED | FLAT, CHUNKED | FRAME):
        o, sT = _run(op, q, k, v, g, beta, S0, frame_tokens=128, flags=flags)
        assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2, flags


def test_strong_decay_takes_the_slow_path(op):
    """Chunk decay below e^-60 (log-gates scaled up): the per-element exp(Gamma_i - Gamma_j) path of the chunk kernel,
    mixed with ordinary chunks in the same clip."""
    q, k, v, g, beta, S0 = make_inputs(2, 5 * 64, 2, 64, 256, seed=33, dtype=torch.bfloat16)
    g = g.clone()
    g[:, 64:192] *= 300.0            # chunks 1 and 2: total decay far below e^-60
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    o, sT = _run(op, q, k, v, g, beta, S0, flags=CHUNKED)
    assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2


def test_cuda_graph_capture_and_two_streams(op):
    """The C-ABI call allocates nothing and never syncs: it can be captured into a CUDA graph and replayed, and
    two problems can run concurrently on two streams."""
    q, k, v, g, beta, S0 = _dev(*make_inputs(3, 4 * 49, 2, 64, 256, seed=34, frame_tokens=49, dtype=torch.bfloat16))
    o_ref, s_ref = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49)
    o = torch.zeros_like(o_ref); sT = torch.zeros_like(s_ref)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        op.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 49)           # warm-up outside capture
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    o.zero_(); sT.zero_()
    with torch.cuda.graph(graph):
        op.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 49)
    for _ in range(2):
        o.zero_(); sT.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(o, o_ref) and torch.equal(sT, s_ref)
    # time segments inside a graph (explicit, and the library's own choice on a shape it cuts): the hand-off scratch becomes
    # allocation / free nodes of the graph
    g2 = torch.cuda.CUDAGraph()
    o.zero_(); sT.zero_()
    with torch.cuda.graph(g2):
        op.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 49, SEG(2))
    for _ in range(3):
        o.zero_(); sT.zero_()
        g2.replay()
        torch.cuda.synchronize()
        assert torch.equal(o, o_ref) and torch.equal(sT, s_ref)
    qb, kb, vb, gb, bb, Sb = _dev(*make_inputs(19, 24 * 64, 8, 64, 256, seed=37, dtype=torch.bfloat16))   # 152 chains: cut
    assert op.plan_segments(qb, kb, vb, gb, bb) > 1
    ob_ref, sb_ref = op.gdr_lkva(qb, kb, vb, gb, bb, None, Sb, True, 0, SEG(1))
    ob, sb2 = torch.zeros_like(ob_ref), torch.zeros_like(sb_ref)
    torch.cuda.synchronize()
    g4 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g4):
        op.gdr_lkva_out(qb, kb, vb, gb, bb, ob, sb2, None, Sb, 0, 0)
    for _ in range(2):
        ob.zero_(); sb2.zero_()
        g4.replay()
        torch.cuda.synchronize()
        assert torch.equal(ob, ob_ref) and torch.equal(sb2, sb_ref)
    # packed variable-length clips inside a graph (unit table, scratch and flags are nodes of the graph)
    lens = [70, 3, 129]
    qp, kp, vp, gp, bp, Sp, cu = _packed(lens, 2, 256, 36)
    qp, kp, vp, gp, bp, Sp = _dev(qp, kp, vp, gp, bp, Sp)
    cud = cu.cuda()
    ov_ref, sv_ref = op.gdr_lkva_varlen(qp, kp, vp, gp, bp, cud, None, Sp, True, CHUNKED)
    torch.cuda.synchronize()
    g3 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g3):
        ov, sv = op.gdr_lkva_varlen(qp, kp, vp, gp, bp, cud, None, Sp, True, CHUNKED)
    for _ in range(2):
        ov.zero_(); sv.zero_()
        g3.replay()
        torch.cuda.synchronize()
        assert torch.equal(ov, ov_ref) and torch.equal(sv, sv_ref)
    # two streams, two different problems, launched back to back
    q2, k2, v2, g2, b2, S2 = _dev(*make_inputs(2, 5 * 64, 3, 64, 128, seed=35, dtype=torch.bfloat16))
    o2_ref, s2_ref = op.gdr_lkva(q2, k2, v2, g2, b2, None, S2, True, 0)
    s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s_a):
        oa, sa = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49)
    with torch.cuda.stream(s_b):
        ob, sb = op.gdr_lkva(q2, k2, v2, g2, b2, None, S2, True, 0)
    torch.cuda.synchronize()
    assert torch.equal(oa, o_ref) and torch.equal(sa, s_ref) and torch.equal(ob, o2_ref) and torch.equal(sb, s2_ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("D", [32, 64, 128, 256])
def test_l2norm_prologue(op, dtype, D):
    """q/k normalisation kernel against x * rsqrt(sum x^2 + eps) in fp32 (the result rounded to the I/O dtype)."""
    g = torch.Generator().manual_seed(40 + D)
    x = (torch.randn(3, 37, 5, D, generator=g) * torch.rand(3, 37, 5, 1, generator=g) * 4).to(dtype)
    ref = (x.float() * torch.rsqrt(x.float().square().sum(-1, keepdim=True) + 1e-6)).to(dtype)
    y = op.l2norm(x.cuda())
    torch.cuda.synchronize()
    assert y.dtype == dtype and y.shape == x.shape
    tol = 1e-6 if dtype == torch.float32 else 2.0 ** -8      # one bf16 ulp of values <= 1
    assert (y.cpu().float() - ref.float()).abs().max().item() <= tol
    z = torch.zeros(4, D, dtype=dtype, device="cuda")          # zero rows stay zero (eps)
    assert torch.equal(op.l2norm(z), z)


def test_chunk_gated_delta_rule_with_qk_l2norm(op):
    """fla's use_qk_l2norm_in_kernel=True: unnormalised q, k in, same result as normalising first."""
    q, k, v, g, beta, S0 = make_inputs(2, 3 * 49, 2, 64, 256, seed=41, frame_tokens=49, dtype=torch.bfloat16)
    q = (q.float() * 3.0).bfloat16(); k = (k.float() * 0.5).bfloat16()
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    o, sT = op.chunk_gated_delta_rule(qd, kd, vd, gd, bd, initial_state=sd, output_final_state=True, use_qk_l2norm_in_kernel=True)
    qn = (q.float() * torch.rsqrt(q.float().square().sum(-1, keepdim=True) + 1e-6)).bfloat16()
    kn = (k.float() * torch.rsqrt(k.float().square().sum(-1, keepdim=True) + 1e-6)).bfloat16()
    o_ref, s_ref = gdr_recurrent_ref(qn, kn, v, g, beta, None, S0)
    assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2


@pytest.mark.parametrize("case", [
    # B, T, H, V, frame_tokens, flags, gate multiplier on tokens 64..191
    (2, 5 * 49, 3, 256, 49, 0, 1.0),
    (2, 4 * 64 + 30, 2, 128, 0, SEG(2), 1.0),
    (2, 5 * 64, 2, 256, 0, 0, 300.0),            # chunks 1-2 take the slow (per-element decay) path
])
def test_qk_l2norm_with_wide_norm_spread(op, case):
    """use_qk_l2norm_in_kernel=True with un-normalised q, k whose row norms spread over several decades: the
    normalisation pass + the op against the oracle on fp32-normalised inputs."""
    B, T, H, V, C, fl, gmul = case
    q, k, v, g, beta, S0 = make_inputs(B, T, H, 64, V, seed=61, frame_tokens=C, correlated=(C == 49), dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(62)
    amp = lambda t: (t.float() * torch.exp(2.3 * torch.randn(B, T, H, 1, generator=gen))).bfloat16()
    q, k = amp(q), amp(k)
    g = g.clone()
    g[:, 64:192] *= gmul
    l2 = lambda t: t.float() * torch.rsqrt((t.float() ** 2).sum(-1, keepdim=True) + 1e-6)
    o_ref, s_ref = gdr_recurrent_ref(l2(q).bfloat16(), l2(k).bfloat16(), v, g, beta, None, S0)
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    o, sT = op.chunk_gated_delta_rule(qd, kd, vd, gd, bd, initial_state=sd, output_final_state=True, frame_tokens=C,
                                      flags=CHUNKED | fl, use_qk_l2norm_in_kernel=True)
    assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2


def test_qk_l2norm_packed_clips(op):
    lens = [200, 64, 1, 333, 17]
    q, k, v, g, beta, S0, cu = _packed(lens, 2, 256, 63)
    q, k = (q.float() * 3.0).bfloat16(), (k.float() * 0.2).bfloat16()
    l2 = lambda t: t.float() * torch.rsqrt((t.float() ** 2).sum(-1, keepdim=True) + 1e-6)
    o_ref, s_ref = gdr_recurrent_varlen_ref(l2(q).bfloat16(), l2(k).bfloat16(), v, g, beta, cu, None, S0)
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    o, sT = op.chunk_gated_delta_rule(qd, kd, vd, gd, bd, initial_state=sd, output_final_state=True, cu_seqlens=cu.cuda(),
                                      use_qk_l2norm_in_kernel=True)
    assert max_rel_err(o.float().cpu(), o_ref) <= 2e-2 and max_rel_err(sT.cpu(), s_ref) <= 2e-2


def test_kat_on_device(op):
    """Orthonormal keys, g=0, beta=1: S = sum k_i v_i^T exactly; reading q=k_j returns scale*v_j."""
    K, V = 64, 64
    keys = torch.eye(K)[None, :, None, :].contiguous()
    vals = torch.arange(K * V, dtype=torch.float32).reshape(1, K, 1, V) / 64.0
    for flags in (RECURRENT, 0):
        o, S = _run(op, keys, keys, vals, torch.zeros(1, K, 1), torch.ones(1, K, 1), None, scale=0.5, flags=flags)
        assert max_rel_err(S[0, 0], vals[0, :, 0]) < 1e-3
        assert max_rel_err(o, 0.5 * vals) < 1e-3


def test_errors_are_loud(op):
    q, k, v, g, beta, S0 = _dev(*make_inputs(1, 16, 1, 64, 64, seed=1))
    with pytest.raises(RuntimeError, match="unsupported shape"):
        op.gdr_lkva(q[..., :48].contiguous(), k[..., :48].contiguous(), v, g, beta)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        op.gdr_lkva(q, k, v, g, beta, frame_tokens=5)
    with pytest.raises(ValueError):
        op.gdr_lkva(q, k, v, g, beta, None, S0[..., :32])


def test_host_pipeline_matches_device_call(op):
    from gdkvm_b200.host import gdr_lkva_host
    q, k, v, g, beta, S0 = make_inputs(5, 3 * 49, 2, 64, 256, seed=12, frame_tokens=49, dtype=torch.bfloat16)
    pin = lambda t: t.pin_memory()
    o_h, s_h = gdr_lkva_host(pin(q), pin(k), pin(v), pin(g), pin(beta), None, pin(S0), 49, clips_per_group=2)
    o, sT = _run(op, q, k, v, g, beta, S0, frame_tokens=49)
    assert torch.equal(o_h.float(), o) and torch.equal(s_h, sT)


# ---------------- full-size (BASELINE configs[1]) properties ----------------

@pytest.mark.parametrize("shape", [
    # B, T, H, V, frame_tokens, flags
    (3, 9 * 64 + 17, 2, 256, 0, 0),            # flat tiling, ragged tail in the last segment
    (2, 7 * 49, 3, 256, 49, FRAME),            # one padded chunk per frame
    (2, 3 * 128, 2, 128, 128, 0),              # V = 128 (one state warpgroup), 2 sub-chunks per frame
    (40, 6 * 64, 8, 256, 0, 0),                # 320 chains x up to 6 segments: several waves of units, real waiting
])
def test_time_segments_are_bit_identical(op, shape):
    """Chains cut into time segments (separate work units, fp32 state handed over through global memory) give
    exactly the bits of the uncut chains: readout, final state, with and without an initial state."""
    B, T, H, V, C, fl = shape
    q, k, v, g, beta, S0 = _dev(*make_inputs(B, T, H, 64, V, seed=41, frame_tokens=C, dtype=torch.bfloat16))
    for s0 in (S0, None):
        o1, s1 = op.gdr_lkva(q, k, v, g, beta, None, s0, True, C, CHUNKED | fl | SEG(1))
        for n in (2, 3, 5, 15):
            o, sT = op.gdr_lkva(q, k, v, g, beta, None, s0, True, C, CHUNKED | fl | SEG(n))
            assert torch.equal(o, o1) and torch.equal(sT, s1), (n, s0 is None)
    o1, _ = op.gdr_lkva(q, k, v, g, beta, None, S0, True, C, CHUNKED | fl | SEG(1))
    o, _ = op.gdr_lkva(q, k, v, g, beta, None, S0, False, C, CHUNKED | fl | SEG(4))     # no final state requested
    assert torch.equal(o, o1)
    if B <= 3:
        qc, kc, vc, gc, bc, sc = (t.cpu() for t in (q, k, v, g, beta, S0))
        o_ref, s_ref = gdr_recurrent_ref(qc, kc, vc, gc, bc, None, sc)
        o, sT = op.gdr_lkva(q, k, v, g, beta, None, S0, True, C, CHUNKED | fl | SEG(3))
        assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2


def test_time_segments_with_slow_path_chunks(op):
    q, k, v, g, beta, S0 = make_inputs(2, 6 * 64, 2, 64, 256, seed=42, dtype=torch.bfloat16)
    g = g.clone()
    g[:, 64:256] *= 300.0
    q, k, v, g, beta, S0 = _dev(q, k, v, g, beta, S0)
    o1, s1 = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 0, CHUNKED | SEG(1))
    for n in (2, 3, 6):
        o, sT = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 0, CHUNKED | SEG(n))
        assert torch.equal(o, o1) and torch.equal(sT, s1), n


def _packed(lens, H, V, seed, dtype=torch.bfloat16, K=64):
    """Clips of the given lengths packed back to back: ([1,T,H,*] tensors, cu_seqlens, per-clip views, S0 [N,H,K,V])."""
    T = sum(lens)
    q, k, v, g, beta, _ = make_inputs(1, max(T, 1), H, K, V, seed=seed, dtype=dtype)
    q, k, v, g, beta = (t[:, :T] for t in (q, k, v, g, beta))
    gen = torch.Generator().manual_seed(seed + 1)
    S0 = 0.1 * torch.randn(len(lens), H, K, V, generator=gen)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int64)
    return q, k, v, g, beta, S0, cu


def _varlen_ref(q, k, v, g, beta, S0, cu):
    return gdr_recurrent_varlen_ref(q, k, v, g, beta, cu, None, S0)


@pytest.mark.parametrize("case", [
    # lens, H, V, dtype, flags
    ([200, 64, 1, 333, 128, 17], 2, 256, torch.bfloat16, CHUNKED),          # ragged tails, exact chunks, single token
    ([700, 0, 65, 0, 1290], 3, 128, torch.bfloat16, CHUNKED | SEG(3)),       # empty clips, forced time segments, V = 128
    ([16 * 64, 64, 10 * 64], 2, 256, torch.bfloat16, CHUNKED | SEG(3)),      # 3-chunk target segments: 16 chunks -> 4+4+4+4
    ([5, 9, 130], 2, 40, torch.float32, 0),                                  # fp32 I/O -> recurrent kernel
    ([100, 37, 260], 2, 256, torch.bfloat16, RECURRENT),                     # recurrent kernel, bf16 I/O
    ([64 * 30] * 5 + [64 * 7 + 3] * 6, 8, 256, torch.bfloat16, 0),           # many units, library-chosen segment length
])
@pytest.mark.parametrize("cu_dtype", [torch.int64, torch.int32], ids=["i64", "i32"])
def test_varlen_packed_clips_vs_oracle(op, case, cu_dtype):
    """Packed clips of different lengths (cu_seqlens on the device) against the per-clip CPU oracle: readout of every
    clip, final state of every clip, nothing written outside a clip's rows."""
    lens, H, V, dtype, flags = case
    K = 64 if V != 40 else 32
    q, k, v, g, beta, S0, cu = _packed(lens, H, V, 51, dtype, K)
    o_ref, s_ref = _varlen_ref(q, k, v, g, beta, S0, cu)
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    o, sT = op.gdr_lkva_varlen(qd, kd, vd, gd, bd, cu.to(cu_dtype).cuda(), None, sd, True, flags)
    torch.cuda.synchronize()
    tol = TOL[dtype]
    assert max_rel_err(o.float().cpu(), o_ref) <= tol and max_rel_err(sT.cpu(), s_ref) <= tol
    for n in range(len(lens)):                       # per clip (a short clip must not hide behind a long one's scale)
        a, b = int(cu[n]), int(cu[n + 1])
        if b > a:
            assert max_rel_err(o[:, a:b].float().cpu(), o_ref[:, a:b]) <= tol, n
        assert max_rel_err(sT[n].cpu(), s_ref[n]) <= tol, n


def test_varlen_many_short_clips(op, c_oracle):
    """700 clips of 0..130 tokens (several rounds of the 256-thread table builder, most clips a single ragged chunk)."""
    rng = np.random.default_rng(54)
    lens = [int(x) for x in rng.integers(0, 131, size=700)]
    q, k, v, g, beta, S0, cu = _packed(lens, 2, 256, 55)
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    o, sT = op.gdr_lkva_varlen(qd, kd, vd, gd, bd, cu.cuda(), None, sd, True, 0)
    o_r, s_r = op.gdr_lkva_varlen(qd, kd, vd, gd, bd, cu.cuda(), None, sd, True, RECURRENT)     # fp32-exact CUDA-core path
    torch.cuda.synchronize()
    assert max_rel_err(o, o_r) <= 2e-2 and max_rel_err(sT, s_r) <= 2e-2
    for n in (0, 1, 2, 350, 698, 699):                                    # spot-check clips against the CPU oracle
        a, b = int(cu[n]), int(cu[n + 1])
        if b == a:
            assert torch.equal(sT[n].cpu(), S0[n])
            continue
        o_ref, s_ref = c_oracle.gdr_recurrent_c(q[:, a:b], k[:, a:b], v[:, a:b], g[:, a:b], beta[:, a:b], None, S0[n:n + 1])
        assert max_rel_err(o[:, a:b].float().cpu(), o_ref) <= 2e-2 and max_rel_err(sT[n:n + 1].cpu(), s_ref) <= 2e-2, n


def test_v64_on_the_chunk_kernel(op):
    """d_v = 64 per head: tcgen05 path (one value block per head, nothing written into the neighbouring head), frame-aligned
    and flat tiling, time segments, packed clips -- against the fp32-exact recurrent kernel and the oracle."""
    q, k, v, g, beta, S0 = make_inputs(2, 6 * 64, 3, 64, 64, seed=71, frame_tokens=64, dtype=torch.bfloat16)
    assert op.plan(q.cuda(), k.cuda(), v.cuda(), g.cuda(), beta.cuda()) == 1
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    for fl in (CHUNKED, CHUNKED | FRAME, CHUNKED | SEG(3)):
        o, sT = _run(op, q, k, v, g, beta, S0, frame_tokens=64, flags=fl)
        assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2, fl
    lens = [100, 64, 7, 200]
    qp, kp, vp, gp, bp, Sp, cu = _packed(lens, 3, 64, 72)
    ov_ref, sv_ref = gdr_recurrent_varlen_ref(qp, kp, vp, gp, bp, cu, None, Sp)
    qd, kd, vd, gd, bd, sd = _dev(qp, kp, vp, gp, bp, Sp)
    ov, sv = op.gdr_lkva_varlen(qd, kd, vd, gd, bd, cu.cuda(), None, sd, True, CHUNKED)
    assert max_rel_err(ov.float().cpu(), ov_ref) <= 2e-2 and max_rel_err(sv.cpu(), sv_ref) <= 2e-2


def test_varlen_equals_batched_call_bit_for_bit(op):
    """Equal-length clips: the packed call must reproduce the batched call exactly (same chunking, same arithmetic),
    and chunk_gated_delta_rule(cu_seqlens=...) is the same entry point."""
    B, T, H, V = 5, 6 * 64 + 21, 4, 256
    q, k, v, g, beta, S0 = _dev(*make_inputs(B, T, H, 64, V, seed=52, dtype=torch.bfloat16))
    o_b, s_b = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 0, CHUNKED)
    pk = lambda t: t.reshape(1, B * T, *t.shape[2:])
    cu = torch.arange(B + 1, device="cuda", dtype=torch.int32) * T
    o_p, s_p = op.gdr_lkva_varlen(pk(q), pk(k), pk(v), pk(g), pk(beta), cu, None, S0, True, CHUNKED)
    assert torch.equal(o_p.reshape(B, T, H, V), o_b) and torch.equal(s_p, s_b)
    o_f, s_f = op.chunk_gated_delta_rule(pk(q), pk(k), pk(v), pk(g), pk(beta), initial_state=S0, output_final_state=True,
                                         cu_seqlens=cu)
    assert torch.equal(o_f, o_p) and torch.equal(s_f, s_p)
    o_n, s_n = op.gdr_lkva_varlen(pk(q), pk(k), pk(v), pk(g), pk(beta), cu, None, None, False, CHUNKED | SEG(2))
    o_b0, _ = op.gdr_lkva(q, k, v, g, beta, None, None, True, 0, CHUNKED)
    assert s_n is None and torch.equal(o_n.reshape(B, T, H, V), o_b0)


def test_varlen_rows_outside_every_clip_are_untouched(op):
    """Caller-owned, NaN-poisoned readout buffer with rows that belong to no clip (a gap before the first clip and behind
    the last one): the packed call -- TMA stores for whole chunks, row-wise stores for a clip's last chunk -- must write
    every row of every clip and not one bit elsewhere, on the tcgen05 and on the recurrent path."""
    lens = [70, 3, 129, 64, 1]
    lead, trail = 5, 9
    T = lead + sum(lens) + trail
    q, k, v, g, beta, _ = make_inputs(1, T, 2, 64, 256, seed=53, dtype=torch.bfloat16)
    S0 = 0.1 * torch.randn(len(lens), 2, 64, 256, generator=torch.Generator().manual_seed(54))
    cu = torch.tensor([lead] + [lead + int(x) for x in np.cumsum(lens)], dtype=torch.int64)
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    for flags in (CHUNKED, CHUNKED | SEG(2), RECURRENT):
        o = torch.full((1, T, 2, 256), float("nan"), dtype=torch.bfloat16, device="cuda")
        sT = torch.full((len(lens), 2, 64, 256), float("nan"), device="cuda")
        op.gdr_lkva_varlen_out(qd, kd, vd, gd, bd, cu.cuda(), o, sT, None, sd, flags)
        torch.cuda.synchronize()
        assert bool(torch.isnan(o[:, :lead]).all()) and bool(torch.isnan(o[:, T - trail:]).all()), flags
        assert not bool(torch.isnan(o[:, lead:T - trail]).any()) and not bool(torch.isnan(sT).any()), flags
        for n in range(len(lens)):          # every clip alone (its own launch) gives the same rows
            a, b = int(cu[n]), int(cu[n + 1])
            o_n, s_n = gdr_recurrent_ref(q[:, a:b], k[:, a:b], v[:, a:b], g[:, a:b], beta[:, a:b], None, S0[n:n + 1])
            assert max_rel_err(o[:, a:b], o_n) <= 2e-2 and max_rel_err(sT[n:n + 1], s_n) <= 2e-2, (flags, n)


def test_varlen_all_clips_empty_pass_their_state_through(op):
    q, k, v, g, beta, _ = _dev(*make_inputs(1, 0, 2, 64, 256, seed=55, dtype=torch.bfloat16))
    S0 = torch.randn(3, 2, 64, 256, device="cuda")
    cu = torch.zeros(4, dtype=torch.int32, device="cuda")
    o, sT = op.gdr_lkva_varlen(q, k, v, g, beta, cu, None, S0, True, 0)
    assert o.shape == (1, 0, 2, 256) and torch.equal(sT, S0)
    _, sT0 = op.gdr_lkva_varlen(q, k, v, g, beta, cu, None, None, True, 0)
    assert torch.equal(sT0, torch.zeros_like(S0))


def test_mixed_devices_are_rejected_before_the_launch(op):
    """A host tensor (or another GPU's) among the arguments would hand the kernel a pointer it cannot dereference."""
    q, k, v, g, beta, S0 = make_inputs(1, 64, 1, 64, 64, seed=2, dtype=torch.bfloat16)
    qd, kd, vd, gd, bd, sd = _dev(q, k, v, g, beta, S0)
    for bad in ("g", "beta", "k", "initial_state"):
        args = dict(q=qd, k=kd, v=vd, g=gd, beta=bd, initial_state=sd)
        args[bad] = dict(q=q, k=k, v=v, g=g, beta=beta, initial_state=S0)[bad]
        with pytest.raises((ValueError, RuntimeError)):
            op.gdr_lkva(args["q"], args["k"], args["v"], args["g"], args["beta"], None, args["initial_state"])
    with pytest.raises(ValueError):
        op.gdr_lkva_out(qd, kd, vd, gd, bd, torch.empty(1, 64, 1, 64, dtype=torch.bfloat16), None)
    cu = torch.tensor([0, 64], dtype=torch.int32)
    with pytest.raises((ValueError, RuntimeError)):
        op.gdr_lkva_varlen(qd, kd, vd, gd, bd, cu)                   # cu_seqlens on the host
    torch.cuda.synchronize()                                        # the context is still alive
    o, _ = op.gdr_lkva(qd, kd, vd, gd, bd, None, sd)
    assert bool(torch.isfinite(o.float()).all())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_one_process(op):
    """The tcgen05 kernel's shared-memory opt-in is per device: cuda:1 after cuda:0 in the same process."""
    q, k, v, g, beta, S0 = make_inputs(2, 5 * 64, 2, 64, 256, seed=91, dtype=torch.bfloat16)
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        t = [x.to(dev) for x in (q, k, v, g, beta, S0)]
        o, sT = op.gdr_lkva(*t[:5], None, t[5], True, 0, CHUNKED | SEG(2))
        torch.cuda.synchronize(dev)
        outs.append((o.cpu(), sT.cpu()))
    assert all(torch.equal(outs[0][0], x[0]) and torch.equal(outs[0][1], x[1]) for x in outs[1:])
    # the same for the training forward, the backward kernel and both tile shapes of the projection kernel
    gen = torch.Generator().manual_seed(92)
    do, dsT = torch.randn(2, 5 * 64, 2, 256, generator=gen).bfloat16(), torch.randn(2, 2, 64, 256, generator=gen)
    x = torch.randn(300, 128, generator=gen).bfloat16()
    w = (torch.randn(2 * (128 + 64) + 4, 128, generator=gen) / 11).bfloat16()
    res = []
    for dev in ("cuda:1", "cuda:0", "cuda:1"):
        t = [y.to(dev) for y in (q, k, v, g, beta, S0, do, dsT, x, w)]
        o, sT, cs = torch.ops.gdkvm.gdr_lkva_train(*t[:5], None, t[5], 0)
        grads = torch.ops.gdkvm.gdr_lkva_bwd(*t[:5], cs, t[6], t[7], 0.125, True, None, 0)
        proj = op.qkvgb_project(t[8], t[9], None, 2, 64, 64)
        torch.cuda.synchronize(dev)
        res.append([y.cpu() for y in (o, sT, *grads, *proj)])
    for other in res[1:]:
        assert all(torch.equal(a, b) for a, b in zip(res[0], other))


def test_unforced_fallback_warns_once_with_the_reason(op):
    q, k, v, g, beta, S0 = _dev(*make_inputs(1, 40, 1, 64, 64, seed=3))         # fp32 I/O: not the tcgen05 kernel
    assert "fp32" in op.plan_reason(q, k, v, g, beta) and op.plan(q, k, v, g, beta) == 0
    assert op.plan_reason(q.bfloat16(), k.bfloat16(), v.bfloat16(), g, beta) == ""
    from gdkvm_b200 import ops
    ops._warned_fallback.clear()
    with pytest.warns(RuntimeWarning, match="fp32"):
        op.gdr_lkva(q, k, v, g, beta, None, S0)
    import warnings as w
    with w.catch_warnings():
        w.simplefilter("error")
        op.gdr_lkva(q, k, v, g, beta, None, S0)                                   # second call: silent
        op.gdr_lkva(q, k, v, g, beta, None, S0, True, 0, RECURRENT)              # forced: never warns


@pytest.mark.parametrize("norm", [1.0, 1.25, 1.5, 2.0, 4.0])
def test_key_norm_envelope_of_the_fp16_solve(op, norm):
    """Un-normalised, strongly correlated keys (one base key per 64-token frame + 0.3 noise, |k| = norm): the intra-chunk
    system (I + A), A_ij = beta_i k_i.k_j, is solved in fp16 on the tcgen05 path.  Inside the stability envelope of the
    delta rule itself -- beta |k|^2 <= 2, here beta is drawn in (0, 1.8 / norm^2) -- the chunk kernel must hold the bf16
    tolerance against the fp32 oracle; check_inputs() is the loud check of that precondition."""
    C = 64
    q, k, v, g, beta, S0 = make_inputs(2, 6 * C, 2, 64, 256, seed=81, frame_tokens=C, correlated=True, dtype=torch.float32)
    k = (k * norm).bfloat16()
    q, v = q.bfloat16(), v.bfloat16()
    beta = beta * min(1.0, 1.8 / norm ** 2)
    assert op.check_inputs(k.cuda(), beta.cuda()) <= 2.0
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    o, sT = _run(op, q, k, v, g, beta, S0, frame_tokens=C, flags=CHUNKED)
    eo, es = max_rel_err(o, o_ref), max_rel_err(sT, s_ref)
    print(f"|k| = {norm}: readout max-rel {eo:.2e} rms {rms_rel_err(o, o_ref):.2e}, final state max-rel {es:.2e}")
    assert eo <= 2e-2 and es <= 2e-2
    with pytest.raises(ValueError, match="unstable"):
        op.check_inputs((k.float() * 2.0).cuda(), torch.ones_like(beta).cuda())


@pytest.fixture(scope="module")
def echonet_batch():
    """configs[1]: 64 clips x 128 frames x 49 tokens, 8 heads, K=64, V=256, bf16 -- built on the GPU."""
    g = torch.Generator(device="cuda").manual_seed(1234)
    B, T, H, K, V = 64, 128 * 49, 8, 64, 256
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda", dtype=torch.float32)
    l2 = lambda x: torch.nn.functional.normalize(x, dim=-1)
    q = l2(rn(B, T, H, K)).bfloat16()
    k = l2(rn(B, T, H, K)).bfloat16()
    v = rn(B, T, H, V).bfloat16()
    beta = torch.sigmoid(rn(B, T, H))
    gate = torch.nn.functional.logsigmoid(rn(B, T, H) + 4.0)
    S0 = 0.1 * rn(B, H, K, V)
    return q, k, v, gate, beta, S0


def test_full_size_sample_vs_oracle(op, echonet_batch, c_oracle):
    """Whole configs[1] batch on the GPU; 2 clips x 8 heads checked against the C oracle."""
    q, k, v, g, beta, S0 = echonet_batch
    o, sT = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49)
    torch.cuda.synchronize()
    for b in (0, 63):
        sl = slice(b, b + 1)
        o_ref, s_ref = c_oracle.gdr_recurrent_c(q[sl].cpu(), k[sl].cpu(), v[sl].cpu(), g[sl].cpu(),
                                                beta[sl].cpu(), None, S0[sl].cpu())
        assert max_rel_err(o[sl], o_ref) <= 2e-2 and max_rel_err(sT[sl], s_ref) <= 2e-2


def test_full_size_linearity(op, echonet_batch):
    """(v, S0) -> (o, S_T) is linear; scaling by 2 is exact in binary floating point, so the
    full-size outputs must scale bit-exactly -- a size-independent check of every chain."""
    q, k, v, g, beta, S0 = echonet_batch
    o, sT = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49)
    o2, sT2 = op.gdr_lkva(q, k, v * 2, g, beta, None, S0 * 2, True, 49)
    assert torch.equal(o2.float(), o.float() * 2) and torch.equal(sT2, sT * 2)


def test_full_size_state_carry(op, echonet_batch):
    q, k, v, g, beta, S0 = echonet_batch
    o, sT = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49)
    cut = 64 * 49
    oa, sa = op.gdr_lkva(q[:, :cut], k[:, :cut], v[:, :cut], g[:, :cut], beta[:, :cut], None, S0, True, 49)
    ob, sb = op.gdr_lkva(q[:, cut:], k[:, cut:], v[:, cut:], g[:, cut:], beta[:, cut:], None, sa, True, 49)
    assert torch.equal(oa, o[:, :cut]) and torch.equal(ob, o[:, cut:]) and torch.equal(sb, sT)


def test_full_size_segments_agree(op, echonet_batch):
    """configs[1]: the library's own choice of time segments vs uncut chains vs 4 segments, bit for bit."""
    q, k, v, g, beta, S0 = echonet_batch
    o, sT = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49)
    for n in (1, 4):
        o_n, s_n = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49, SEG(n))
        assert torch.equal(o_n, o) and torch.equal(s_n, sT), n


def test_full_size_paths_agree(op, echonet_batch):
    """Recurrent fp32 kernel vs the default path over the whole batch (chunk-size invariance)."""
    q, k, v, g, beta, S0 = echonet_batch
    o, sT = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49)
    o_r, s_r = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 49, RECURRENT)
    assert max_rel_err(o, o_r) <= 2e-2 and max_rel_err(sT, s_r) <= 2e-2


def _device_batch(B, T, H, K, V, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda", dtype=torch.float32)
    l2 = lambda x: torch.nn.functional.normalize(x, dim=-1)
    q = l2(rn(B, T, H, K)).bfloat16()
    k = l2(rn(B, T, H, K)).bfloat16()
    v = rn(B, T, H, V).bfloat16()
    return q, k, v, torch.nn.functional.logsigmoid(rn(B, T, H) + 4.0), torch.sigmoid(rn(B, T, H)), 0.1 * rn(B, H, K, V)


def _check_clips(o, sT, batch, clips, c_oracle, C, what):
    q, k, v, g, beta, S0 = batch
    worst = 0.0
    for b in clips:
        sl = slice(b, b + 1)
        o_ref, s_ref = c_oracle.gdr_recurrent_c(q[sl].cpu(), k[sl].cpu(), v[sl].cpu(), g[sl].cpu(), beta[sl].cpu(), None, S0[sl].cpu())
        _assert_per_chain(o[sl].float().cpu(), o_ref, sT[sl].cpu(), s_ref, 2e-2, (what, b))
        worst = max(worst, float(per_frame_max_rel(o[sl].float().cpu(), o_ref, C).max()))
    return worst


def test_full_size_camus_vs_oracle(op, c_oracle):
    """BASELINE configs[2]: 32 clips x 20 frames x 1024 tokens (256 x 256 frames, stride 8), 8 heads -- the whole batch on the
    GPU, 2 clips x 8 heads against the C oracle, per chain."""
    batch = _device_batch(32, 20 * 1024, 8, 64, 256, 2222)
    q, k, v, g, beta, S0 = batch
    o, sT = op.gdr_lkva(q, k, v, g, beta, None, S0, True, 1024)
    torch.cuda.synchronize()
    assert op.plan(q, k, v, g, beta, frame_tokens=1024) == 1
    worst = _check_clips(o, sT, batch, (0, 31), c_oracle, 1024, "camus")
    print(f"configs[2] full size: worst per-frame readout max-rel {worst:.2e}")
    o2, sT2 = op.gdr_lkva(q, k, v * 2, g, beta, None, S0 * 2, True, 1024)               # linearity: every chain, bit-exact
    assert torch.equal(o2.float(), o.float() * 2) and torch.equal(sT2, sT * 2)


def test_full_size_long_clips_two_chained_calls_vs_oracle(op, c_oracle):
    """BASELINE configs[3] (one GPU's share): 64 clips x 256 frames x 49 tokens, 8 heads, run as TWO chained calls (frames
    0-127, then 128-255 from the first call's final state): 2 clips x 8 heads against the C oracle run over the whole
    256-frame clip, per chain, with the error-growth curve over the frames."""
    C, F = 49, 256
    batch = _device_batch(64, F * C, 8, 64, 256, 3333)
    q, k, v, g, beta, S0 = batch
    cut = (F // 2) * C
    oa, sa = op.gdr_lkva(q[:, :cut], k[:, :cut], v[:, :cut], g[:, :cut], beta[:, :cut], None, S0, True, C)
    ob, sb = op.gdr_lkva(q[:, cut:], k[:, cut:], v[:, cut:], g[:, cut:], beta[:, cut:], None, sa, True, C)
    torch.cuda.synchronize()
    o = torch.cat([oa, ob], 1)
    worst = 0.0
    for b in (0, 63):
        sl = slice(b, b + 1)
        o_ref, s_ref = c_oracle.gdr_recurrent_c(q[sl].cpu(), k[sl].cpu(), v[sl].cpu(), g[sl].cpu(), beta[sl].cpu(), None, S0[sl].cpu())
        _assert_per_chain(o[sl].float().cpu(), o_ref, sb[sl].cpu(), s_ref, 2e-2, ("long_clip", b))
        curve = per_frame_max_rel(o[sl].float().cpu(), o_ref, C)
        worst = max(worst, float(curve.max()))
        # no drift: the last 32 frames are no worse than twice the first 32 (the state is a contraction, alpha < 1)
        assert float(curve[-32:].max()) <= max(2.0 * float(curve[:32].max()), 1e-2), curve
    print(f"configs[3] full size, two chained calls: worst per-frame readout max-rel {worst:.2e}")
    o1, s1 = op.gdr_lkva(q, k, v, g, beta, None, S0, True, C, FRAME)                      # one call, frame-aligned chunks
    oa2, sa2 = op.gdr_lkva(q[:, :cut], k[:, :cut], v[:, :cut], g[:, :cut], beta[:, :cut], None, S0, True, C, FRAME)
    ob2, sb2 = op.gdr_lkva(q[:, cut:], k[:, cut:], v[:, cut:], g[:, cut:], beta[:, cut:], None, sa2, True, C, FRAME)
    assert torch.equal(torch.cat([oa2, ob2], 1), o1) and torch.equal(sb2, s1)            # same chunking: bit-identical


def test_recurrent_kernel_two_columns_per_thread(op):
    """With >= 444 CTAs and d_v > 128 the fp32 kernel gives every thread two value columns (x and x + 128): a ragged second
    column block (d_v = 192), a full one (d_v = 256), fp32 and bf16 I/O, against the oracle -- and equal to the one-column
    layout of the same arithmetic (a batch too small for the switch) bit for bit."""
    for V, dtype in ((192, torch.float32), (256, torch.float32), (256, torch.bfloat16)):
        q, k, v, g, beta, S0 = make_inputs(60, 37, 8, 64, V, seed=900 + V, dtype=dtype)
        o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
        o, sT = _run(op, q, k, v, g, beta, S0, flags=RECURRENT)
        assert max_rel_err(o, o_ref) <= (1e-5 if dtype == torch.float32 else TOL[dtype]) and max_rel_err(sT, s_ref) <= 1e-5
        o1, s1 = _run(op, q[:3], k[:3], v[:3], g[:3], beta[:3], S0[:3], flags=RECURRENT)          # 24 chains: one column per thread
        assert torch.equal(o1, o[:3]) and torch.equal(s1, sT[:3])


def test_mixed_plan_is_bit_identical_to_the_uniform_plans(op):
    """A batch that fills more than one wave of SMs is scheduled with whole waves of uncut chains and only the clips of the last
    wave cut (gdkvm_gdr_plan_units): same bits as every uniform plan, with and without an initial state, in a CUDA graph too; a
    strided batch view or an explicit segment count keeps the uniform plan."""
    B, T, H, V = 40, 20 * 64 + 9, 8, 128                  # 320 chains on 148 SMs: 18 clips uncut, 22 clips in two segments
    q, k, v, g, beta, S0 = _dev(*make_inputs(B, T, H, 64, V, seed=77, dtype=torch.bfloat16))
    plan = op.plan_units(q, k, v, g, beta)
    assert plan["mixed"] and plan["uncut_clips"] + plan["cut_clips"] == B and plan["cut_clips"] > 0 and plan["segments"] > 1
    assert not op.plan_units(q, k, v, g, beta, flags=SEG(2))["mixed"] and not op.plan_units(q[::2], k[::2], v[::2], g[::2], beta[::2])["mixed"]
    for s0 in (S0, None):
        o_m, s_m = op.gdr_lkva(q, k, v, g, beta, None, s0, True, 0, CHUNKED)
        for n in (1, 2, 3):
            o_u, s_u = op.gdr_lkva(q, k, v, g, beta, None, s0, True, 0, CHUNKED | SEG(n))
            assert torch.equal(o_m, o_u) and torch.equal(s_m, s_u), n
    # frames of whole 64-token chunks (the CAMUS case): the same chunks, the same plan, the same bits
    fq, fk, fv, fg, fb = (t[:, :1280].contiguous() for t in (q, k, v, g, beta))
    assert op.plan_units(fq, fk, fv, fg, fb, frame_tokens=128)["mixed"]
    o_f, s_f = op.gdr_lkva(fq, fk, fv, fg, fb, None, None, True, 128, CHUNKED)
    o_g, s_g = op.gdr_lkva(fq, fk, fv, fg, fb, None, None, True, 128, CHUNKED | SEG(1))
    assert torch.equal(o_f, o_g) and torch.equal(s_f, s_g)
    o = torch.zeros_like(o_m); sT = torch.zeros_like(s_m)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        op.gdr_lkva_out(q, k, v, g, beta, o, sT, None, None, 0, CHUNKED)
    for _ in range(2):
        o.zero_(); sT.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(o, o_m) and torch.equal(sT, s_m)
    o_ref, s_ref = gdr_recurrent_ref(*[t[-2:].cpu() for t in (q, k, v, g, beta)], None, None)       # two of the cut clips against the oracle
    assert max_rel_err(o_m[-2:].float().cpu(), o_ref) <= TOL[torch.bfloat16] and max_rel_err(s_m[-2:].cpu(), s_ref) <= TOL[torch.bfloat16]
