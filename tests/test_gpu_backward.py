"""GPU tests of the backward pass: csrc/gdr_bwd_sm100.cu through torch autograd and through
the C ABI, against reverse-mode differentiation of the token recurrence in float64 (oracle.gdr_ref.gdr_backward_ref) on
the same bf16-rounded inputs.  Tolerance: max|a-b| / max|b| <= 2e-2 per gradient tensor (north_star's bf16 I/O bound)."""
import pytest
import torch

from oracle.gdr_ref import gdr_backward_ref, gdr_recurrent_ref, make_inputs, max_rel_err, rms_rel_err

pytestmark = pytest.mark.gpu
NAMES = ("dq", "dk", "dv", "dg", "dbeta", "dS0")
CHUNKED, FLAT = 0x2, 0x4


@pytest.fixture(scope="module")
def op(built_lib):
    assert torch.cuda.is_available(), "-m gpu tests need a B200"
    import gdkvm_b200
    return gdkvm_b200


def _case(B, T, H, V, seed, C=0, corr=False, with_s0=True, with_dsT=True):
    q, k, v, g, beta, S0 = make_inputs(B, T, H, 64, V, seed=seed, frame_tokens=C, correlated=corr, dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(seed + 1000)
    do = torch.randn(B, T, H, V, generator=gen).bfloat16()
    dsT = torch.randn(B, H, 64, V, generator=gen) if with_dsT else None
    return q, k, v, g, beta, (S0 if with_s0 else None), do, dsT


def _grads_on_device(op, q, k, v, g, beta, S0, do, dsT):
    leaf = lambda x: x.cuda().requires_grad_(True) if x is not None else None
    qd, kd, vd, gd, bd, sd = map(leaf, (q, k, v, g, beta, S0))
    o, sT = op.gdr_lkva(qd, kd, vd, gd, bd, None, sd, True)
    assert o.requires_grad and sT.requires_grad
    loss = (o.float() * do.cuda().float()).sum()
    if dsT is not None:
        loss = loss + (sT * dsT.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    return [x.grad.float().cpu() if x is not None else None for x in (qd, kd, vd, gd, bd, sd)], o.detach(), sT.detach()


@pytest.mark.parametrize("case", [
    # B, T, H, V, frame_tokens, correlated, with S0, with dsT
    (1, 64, 1, 128, 0, False, True, True),              # one full chunk, one value half
    (2, 3 * 64 + 10, 2, 256, 0, False, True, True),     # ragged tail, two value halves
    (2, 5 * 49, 3, 256, 49, True, True, False),         # EchoNet-shaped frames, correlated keys, no final-state cotangent
    (1, 37, 2, 128, 0, False, False, True),             # less than a chunk, zero initial state
    (1, 8 * 64, 2, 256, 0, False, True, True),          # eight chunks: the state cotangent carried a long way
    (2, 3 * 64 + 5, 3, 64, 0, True, True, True),        # d_v = 64: half-empty value tiles
])
def test_backward_vs_float64_autograd(op, case):
    B, T, H, V, C, corr, w_s0, w_dsT = case
    q, k, v, g, beta, S0, do, dsT = _case(B, T, H, V, 200 + T, C, corr, w_s0, w_dsT)
    ref = gdr_backward_ref(q, k, v, g, beta, do, dsT, None, S0)
    got, o, sT = _grads_on_device(op, q, k, v, g, beta, S0, do, dsT)
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2          # the training forward itself
    for name, a, b in zip(NAMES, got, ref):
        if a is None:
            continue
        e, r = max_rel_err(a, b.float()), rms_rel_err(a, b.float())
        print(f"{case}: {name} max-rel {e:.2e} rms-rel {r:.2e}")
        assert e <= 2e-2 and r <= 1e-2, (name, e, r)


@pytest.mark.parametrize("cu_dtype", [torch.int32, torch.int64], ids=["i32", "i64"])
def test_backward_of_packed_variable_length_clips(op, cu_dtype):
    """cu_seqlens call, differentiated: ragged clips, an empty clip, a single token, forced time segments in the forward --
    every gradient of every clip against float64 autograd over that clip alone."""
    lens = [200, 0, 64, 1, 333, 17]
    H, V, T = 2, 256, sum(lens)
    q, k, v, g, beta, _, do, _ = _case(1, T, H, V, 401)
    gen = torch.Generator().manual_seed(402)
    S0 = 0.1 * torch.randn(len(lens), H, 64, V, generator=gen)
    dsT = torch.randn(len(lens), H, 64, V, generator=gen)
    cu = torch.tensor([0] + [sum(lens[:i + 1]) for i in range(len(lens))], dtype=cu_dtype)
    leaf = lambda x: x.cuda().requires_grad_(True)
    qd, kd, vd, gd, bd, sd = map(leaf, (q, k, v, g, beta, S0))
    o, sT = op.gdr_lkva_varlen(qd, kd, vd, gd, bd, cu.cuda(), None, sd, True, (3 & 0xF) << 8)
    ((o.float() * do.cuda().float()).sum() + (sT * dsT.cuda()).sum()).backward()
    torch.cuda.synchronize()
    got = [x.grad.float().cpu() for x in (qd, kd, vd, gd, bd, sd)]
    for n, L in enumerate(lens):
        a, b = int(cu[n]), int(cu[n + 1])
        if L == 0:
            assert torch.equal(got[5][n], dsT[n])                   # an empty clip passes the state cotangent through
            continue
        ref = gdr_backward_ref(q[:, a:b], k[:, a:b], v[:, a:b], g[:, a:b], beta[:, a:b], do[:, a:b], dsT[n:n + 1], None, S0[n:n + 1])
        for name, x, r in zip(NAMES, got, ref):
            xx = x[n:n + 1] if name == "dS0" else x[:, a:b]
            assert max_rel_err(xx, r.float()) <= 2e-2, (n, name)


def test_backward_with_strong_decay_and_bf16_gates(op):
    """Chunks whose total decay is far below e^-60 (the forward kernel's slow path; exp(Gamma) underflows to zero inside the
    chunk) mixed with ordinary ones, and gates / beta given in bf16 (their gradients come back in bf16)."""
    q, k, v, g, beta, S0, do, dsT = _case(2, 5 * 64, 2, 256, 601)
    g = g.clone()
    g[:, 64:192] *= 300.0
    ref = gdr_backward_ref(q, k, v, g, beta, do, dsT, None, S0)
    got, o, sT = _grads_on_device(op, q, k, v, g, beta, S0, do, dsT)
    for name, a, b in zip(NAMES, got, ref):
        assert bool(torch.isfinite(a).all()), name
        assert max_rel_err(a, b.float()) <= 2e-2, (name, max_rel_err(a, b.float()))
    q, k, v, g, beta, S0, do, dsT = _case(1, 130, 2, 128, 602)
    gb, bb = g.bfloat16(), beta.bfloat16()
    ref = gdr_backward_ref(q, k, v, gb, bb, do, dsT, None, S0)
    got, _, _ = _grads_on_device(op, q, k, v, gb, bb, S0, do, dsT)
    for name, a, b in zip(NAMES, got, ref):
        tol = 2e-2 if name not in ("dg", "dbeta") else 3e-2          # + one bf16 rounding of the returned gate gradients
        assert max_rel_err(a, b.float()) <= tol, (name, max_rel_err(a, b.float()))


def test_backward_over_a_long_clip_does_not_drift(op):
    """64 frames x 49 tokens = 49 chunks (the state cotangent is carried through all of them, with time segments forced so
    that it also crosses two hand-offs): every gradient within tolerance, and the error of dq / dk / dv over the FIRST quarter
    of the clip -- the far end of the reverse scan -- no worse than twice that of the last quarter."""
    T = 64 * 49
    q, k, v, g, beta, S0, do, dsT = _case(1, T, 1, 256, 701, 49, True)
    ref = gdr_backward_ref(q, k, v, g, beta, do, dsT, None, S0)
    qd, kd, vd, gd, bd, sd = (x.cuda() for x in (q, k, v, g, beta, S0))
    _, _, cs = torch.ops.gdkvm.gdr_lkva_train(qd, kd, vd, gd, bd, None, sd, 0)
    got = torch.ops.gdkvm.gdr_lkva_bwd(qd, kd, vd, gd, bd, cs, do.cuda(), dsT.cuda(), 0.125, True, None, 3 << 8)
    torch.cuda.synchronize()
    for name, a, b in zip(NAMES, got, ref):
        e = max_rel_err(a, b.float())
        print(f"long clip: {name} max-rel {e:.2e}")
        assert e <= 2e-2, (name, e)
    for name, a, b in zip(NAMES[:3], got[:3], ref[:3]):
        a, b = a.float().cpu(), b.float()
        den = b.abs().max()
        first = float((a[:, :T // 4] - b[:, :T // 4]).abs().max() / den)
        last = float((a[:, -T // 4:] - b[:, -T // 4:]).abs().max() / den)
        assert first <= max(2.0 * last, 5e-3), (name, first, last)


def test_training_forward_is_the_inference_forward(op):
    """gdr_lkva_train = the tcgen05 kernel on flat 64-token chunks, bit for bit, plus the bf16 chunk-start states."""
    q, k, v, g, beta, S0, _, _ = _case(3, 5 * 64 + 7, 2, 256, 301)
    qd, kd, vd, gd, bd, sd = (x.cuda() for x in (q, k, v, g, beta, S0))
    o, sT = op.gdr_lkva(qd, kd, vd, gd, bd, None, sd, True, 0, CHUNKED | FLAT)
    o2, sT2, cs = torch.ops.gdkvm.gdr_lkva_train(qd, kd, vd, gd, bd, None, sd, 0)
    assert torch.equal(o, o2) and torch.equal(sT, sT2)
    assert cs.shape == (3 * 2, 6, 256, 64)
    assert torch.equal(cs[:, 0].float(), sd.reshape(6, 64, 256).transpose(1, 2).bfloat16().float())      # chunk 0 starts at S0
    # chunk c starts where a call over the first 64 c tokens ends
    _, s3 = op.gdr_lkva(qd[:, :192], kd[:, :192], vd[:, :192], gd[:, :192], bd[:, :192], None, sd, True, 0, CHUNKED | FLAT)
    assert max_rel_err(cs[:, 3].float(), s3.reshape(6, 64, 256).transpose(1, 2)) <= 1e-2


def test_backward_time_segments_are_bit_identical(op):
    """The reverse scan cut into time segments (separate work units, dS handed over through global memory in fp32) gives exactly
    the bits of the uncut chains, and the library's own choice on a shape it cuts (152 chains on 148 SMs) equals them too."""
    for (B, T, H, V, seed) in ((3, 9 * 64 + 17, 2, 256, 501), (19, 24 * 64, 8, 128, 502)):
        q, k, v, g, beta, S0, do, dsT = (x.cuda() for x in _case(B, T, H, V, seed))
        _, _, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
        ref = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, True, None, 1 << 8)
        for n in (2, 3, 5, 0):
            got = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, True, None, n << 8)
            for name, a, b in zip(NAMES, got, ref):
                assert torch.equal(a, b), (B, n, name)


def test_backward_and_projection_are_bit_deterministic(op):
    """compute-sanitizer is closed on this pool, so the cheap race detector is determinism: neither kernel uses atomics on data,
    so ten runs of a multi-wave problem (320 chains on 148 SMs, time segments, both value halves; 40 projection row blocks per
    CTA walk) must agree bit for bit -- any unsynchronised shared-memory hand-off shows up as run-to-run noise."""
    q, k, v, g, beta, S0, do, dsT = (x.cuda() for x in _case(40, 6 * 64 + 9, 8, 256, 801))
    _, _, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
    ref = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, True, None, 0)
    for _ in range(9):
        _, _, cs2 = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
        assert torch.equal(cs2, cs)
        got = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs2, do, dsT, 0.125, True, None, 0)
        for name, a, b in zip(NAMES, got, ref):
            assert torch.equal(a, b), name
    gen = torch.Generator(device="cuda").manual_seed(802)
    x = torch.randn(148 * 128 * 3 + 77, 256, generator=gen, device="cuda").bfloat16()
    w = (torch.randn(8 * 384 + 16, 256, generator=gen, device="cuda") / 16).bfloat16()
    bias = torch.randn(8 * 384 + 16, generator=gen, device="cuda")           # staged through shared memory by the producer warp
    for bb in (None, bias):
        pref = op.qkvgb_project(x, w, bb, 8, 64, 256)
        for _ in range(9):
            for a, b in zip(op.qkvgb_project(x, w, bb, 8, 64, 256), pref):
                assert torch.equal(a, b)


def test_training_step_inside_a_cuda_graph(op):
    """Projection -> training forward -> backward captured into ONE CUDA graph and replayed on new data: none of the three
    launches synchronises or allocates outside the stream (the backward's and the segment scheme's scratch become allocation
    nodes), and the replays reproduce the eager results bit for bit."""
    H, V, D = 8, 256, 256
    gen = torch.Generator(device="cuda").manual_seed(901)
    w = (torch.randn(H * (128 + V) + 2 * H, D, generator=gen, device="cuda") / D ** 0.5).bfloat16()
    bias = 0.1 * torch.randn(w.shape[0], generator=gen, device="cuda")
    bias[-2 * H:-H] += 3.0
    xs = [torch.randn(20, 5 * 64 + 11, D, generator=gen, device="cuda").bfloat16() for _ in range(3)]      # 160 chains: cut into segments
    do = torch.randn(20, 5 * 64 + 11, H, V, generator=gen, device="cuda").bfloat16()
    dsT = torch.randn(20, H, 64, V, generator=gen, device="cuda")

    def step(x):
        q, k, v, g, beta = op.qkvgb_project(x, w, bias, H, 64, V)
        o, sT, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, None, 0)
        return (o, sT) + tuple(torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, False, None, 0))

    eager = [[t.clone() for t in step(x) if t is not None] for x in xs]
    xin = xs[0].clone()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = [t for t in step(xin) if t is not None]
    for x, ref in zip(xs[::-1], eager[::-1]):
        xin.copy_(x)
        graph.replay()
        torch.cuda.synchronize()
        for a, b in zip(outs, ref):
            assert torch.equal(a, b)


def test_module_and_alias_are_differentiable(op):
    q, k, v, g, beta, S0, do, dsT = _case(2, 130, 2, 128, 302)
    leaf = lambda x: x.cuda().requires_grad_(True)
    qd, kd, vd, gd, bd, sd = map(leaf, (q, k, v, g, beta, S0))
    mem = op.GDRMemory(frame_tokens=0)
    o, sT = mem(qd, kd, vd, gd, bd, sd)
    ((o.float() * do.cuda().float()).sum() + (sT * dsT.cuda()).sum()).backward()
    g1 = [x.grad.clone() for x in (qd, kd, vd, gd, bd, sd)]
    for x in (qd, kd, vd, gd, bd, sd):
        x.grad = None
    o2, sT2 = op.chunk_gated_delta_rule(qd, kd, vd, gd, bd, initial_state=sd, output_final_state=True)
    ((o2.float() * do.cuda().float()).sum() + (sT2 * dsT.cuda()).sum()).backward()
    for a, x, n in zip(g1, (qd, kd, vd, gd, bd, sd), NAMES):
        tol = 0 if n in ("dq", "dk", "dv", "dS0") else 1e-4          # dg / dbeta are summed with shared-memory atomics
        assert max_rel_err(x.grad.float(), a.float()) <= tol, n
    # only some inputs need gradients; no_grad runs the plain forward
    o3, _ = op.gdr_lkva(qd.detach(), kd.detach(), vd.detach().requires_grad_(True), gd.detach(), bd.detach(), None, None, False)
    o3.float().sum().backward()
    with torch.no_grad():
        o4, _ = op.gdr_lkva(qd, kd, vd, gd, bd, None, sd, True)
    assert not o4.requires_grad


def test_what_cannot_be_differentiated_says_so(op):
    q, k, v, g, beta, S0 = make_inputs(1, 40, 1, 64, 128, seed=5)                 # fp32 I/O
    qd = q.cuda().requires_grad_(True)
    with pytest.raises(NotImplementedError, match="bf16"):
        op.gdr_lkva(qd, k.cuda(), v.cuda(), g.cuda(), beta.cuda())
    qb, kb, vb = (x.cuda().bfloat16() for x in (q, k, v))
    o, _ = torch.ops.gdkvm.gdr_lkva(qb.requires_grad_(True), kb, vb, g.cuda(), beta.cuda())          # the raw inference op
    with pytest.raises(NotImplementedError, match="no backward formula"):
        o.float().sum().backward()
    cu = torch.tensor([0, 40], dtype=torch.int32, device="cuda")
    o, _ = torch.ops.gdkvm.gdr_lkva_varlen(qb, kb, vb, g.cuda(), beta.cuda(), cu)                    # the raw packed inference op
    with pytest.raises(NotImplementedError, match="no backward formula"):
        o.float().sum().backward()
    with pytest.raises(NotImplementedError):                                                         # d_v = 32
        op.gdr_lkva(qb, kb, vb[..., :32].contiguous().requires_grad_(True), g.cuda(), beta.cuda())


def test_l2norm_is_differentiable(op):
    x = torch.randn(3, 17, 2, 64, generator=torch.Generator().manual_seed(7))
    xd = x.cuda().requires_grad_(True)
    w = torch.randn_like(x).cuda()
    (op.l2norm(xd) * w).sum().backward()
    xr = x.double().requires_grad_(True)
    ((xr * torch.rsqrt(xr.square().sum(-1, keepdim=True) + 1e-6)) * w.cpu().double()).sum().backward()
    assert max_rel_err(xd.grad, xr.grad.float()) <= 1e-5


def test_full_size_backward_linearity(op):
    """BASELINE configs[1] through forward + backward: gradients are linear in the cotangents, and doubling is exact in
    binary floating point -- a size-independent check of every chain (dg / dbeta up to the order of their atomics)."""
    g0 = torch.Generator(device="cuda").manual_seed(77)
    B, T, H, K, V = 64, 128 * 49, 8, 64, 256
    rn = lambda *s: torch.randn(*s, generator=g0, device="cuda", dtype=torch.float32)
    l2 = lambda x: torch.nn.functional.normalize(x, dim=-1)
    q, k, v = l2(rn(B, T, H, K)).bfloat16(), l2(rn(B, T, H, K)).bfloat16(), rn(B, T, H, V).bfloat16()
    beta, g = torch.sigmoid(rn(B, T, H)), torch.nn.functional.logsigmoid(rn(B, T, H) + 4.0)
    S0, do, dsT = 0.1 * rn(B, H, K, V), rn(B, T, H, V).bfloat16(), rn(B, H, K, V)
    o, sT, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
    sc = 1.0 / 8.0
    a = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, sc, True)
    b = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do * 2, dsT * 2, sc, True)
    torch.cuda.synchronize()
    for n, x, y in zip(NAMES, a, b):
        assert bool(torch.isfinite(x.float()).all()), n
        if n in ("dg", "dbeta"):
            assert max_rel_err(y, x * 2) <= 1e-4, n
        else:
            assert torch.equal(y.float(), x.float() * 2), n
    # two clips against the float64 oracle would take minutes at this length; one chain, first 6 frames of clip 0 instead
    sl, tt = slice(0, 1), 6 * 49
    ref = gdr_backward_ref(q[sl, :tt, :1].cpu(), k[sl, :tt, :1].cpu(), v[sl, :tt, :1].cpu(), g[sl, :tt, :1].cpu(), beta[sl, :tt, :1].cpu(),
                           do[sl, :tt, :1].cpu(), None, None, S0[sl, :1].cpu())
    leaf = lambda x: x.clone().requires_grad_(True)
    qs, ks, vs = leaf(q[sl, :tt, :1].contiguous()), leaf(k[sl, :tt, :1].contiguous()), leaf(v[sl, :tt, :1].contiguous())
    gs, bs, ss = leaf(g[sl, :tt, :1].contiguous()), leaf(beta[sl, :tt, :1].contiguous()), leaf(S0[sl, :1].contiguous())
    o1, _ = op.gdr_lkva(qs, ks, vs, gs, bs, None, ss, True)
    (o1.float() * do[sl, :tt, :1].float()).sum().backward()
    for n, x, r in zip(NAMES, (qs, ks, vs, gs, bs, ss), ref):
        assert max_rel_err(x.grad.float(), r.float()) <= 2e-2, n
