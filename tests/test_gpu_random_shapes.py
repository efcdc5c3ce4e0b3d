"""Randomised differential tests on the GPU: seeded random problem shapes (clips, tokens, heads, d_v, frame size, chunking flags,
time segments, packed clip lengths) through every tensor-core entry point against the CPU oracle.  The fixed cases elsewhere pin
the shapes BASELINE names; these look for the shape nobody thought of (tails, single tokens, one chain, many tiny clips)."""
import random

import pytest
import torch

from oracle.gdr_ref import (gdr_backward_ref, gdr_recurrent_ref, gdr_recurrent_varlen_ref, make_inputs, max_rel_err)

pytestmark = pytest.mark.gpu
CHUNKED, FLAT, FRAME = 0x2, 0x4, 0x8
TOL = 2e-2


@pytest.fixture(scope="module")
def op(built_lib):
    assert torch.cuda.is_available(), "-m gpu tests need a B200"
    import gdkvm_b200
    return gdkvm_b200


def _shapes(n, seed):
    rnd = random.Random(seed)
    out = []
    for i in range(n):
        B, H, V = rnd.randint(1, 3), rnd.randint(1, 3), rnd.choice([64, 128, 256])
        kind = rnd.choice(["flat", "frames", "tiny", "long"])
        if kind == "tiny":
            T, C = rnd.randint(1, 20), 0
        elif kind == "frames":
            C = rnd.choice([7, 16, 49, 64, 100, 128, 130, 200])
            T = C * rnd.randint(1, 4)
        elif kind == "long":
            T, C = rnd.randint(400, 900), 0
        else:
            T, C = rnd.randint(21, 300), 0
        out.append((B, T, H, V, C, rnd.random() < 0.4, rnd.choice([0, FLAT, FRAME]), rnd.choice([0, 0, 2, 3]), 7000 + i))
    return out


@pytest.mark.parametrize("case", _shapes(28, 20250))
def test_random_shapes_forward(op, case):
    B, T, H, V, C, corr, chunking, nseg, seed = case
    q, k, v, g, beta, S0 = make_inputs(B, T, H, 64, V, seed=seed, frame_tokens=C, correlated=corr, dtype=torch.bfloat16)
    if seed % 3 == 0:
        S0 = None
    o_ref, s_ref = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    dev = lambda t: t.cuda() if t is not None else None
    flags = (chunking if (C or chunking != FRAME) else 0) | (nseg << 8)
    if op.plan(dev(q), dev(k), dev(v), dev(g), dev(beta), frame_tokens=C) == 1:
        flags |= CHUNKED                    # eligible: make sure it is the tensor-core kernel that answers
    o, sT = op.gdr_lkva(dev(q), dev(k), dev(v), dev(g), dev(beta), None, dev(S0), True, C, flags)
    torch.cuda.synchronize()
    assert max_rel_err(o.float().cpu(), o_ref) <= TOL, case
    assert max_rel_err(sT.cpu(), s_ref) <= TOL, case


def _packs(n, seed):
    rnd = random.Random(seed)
    out = []
    for i in range(n):
        ns = rnd.randint(1, 9)
        lens = [rnd.choice([0, 1, rnd.randint(2, 63), 64, rnd.randint(65, 400)]) for _ in range(ns)]
        if sum(lens) == 0:
            lens[0] = 5
        out.append((lens, rnd.randint(1, 3), rnd.choice([64, 128, 256]), rnd.choice([torch.int32, torch.int64]), 8000 + i))
    return out


@pytest.mark.parametrize("case", _packs(10, 777))
def test_random_packed_clips_forward_and_backward(op, case):
    lens, H, V, cu_dtype, seed = case
    T = sum(lens)
    q, k, v, g, beta, _ = make_inputs(1, T, H, 64, V, seed=seed, dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(seed)
    S0 = 0.1 * torch.randn(len(lens), H, 64, V, generator=gen)
    do = torch.randn(1, T, H, V, generator=gen).bfloat16()
    dsT = torch.randn(len(lens), H, 64, V, generator=gen)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=cu_dtype)
    o_ref, s_ref = gdr_recurrent_varlen_ref(q, k, v, g, beta, cu, None, S0)
    leaf = lambda x: x.cuda().requires_grad_(True)
    qd, kd, vd, gd, bd, sd = map(leaf, (q, k, v, g, beta, S0))
    o, sT = op.gdr_lkva_varlen(qd, kd, vd, gd, bd, cu.cuda(), None, sd, True)
    assert max_rel_err(o.detach().float().cpu(), o_ref) <= TOL and max_rel_err(sT.detach().cpu(), s_ref) <= TOL, case
    ((o.float() * do.cuda().float()).sum() + (sT * dsT.cuda()).sum()).backward()
    torch.cuda.synchronize()
    # float64 autograd through the token recurrence, clip by clip
    refs = [torch.zeros_like(x, dtype=torch.float64) for x in (q, k, v, g, beta, S0)]
    for n, L in enumerate(lens):
        a, b = int(cu[n]), int(cu[n + 1])
        if L == 0:
            refs[5][n] = dsT[n].double()
            continue
        gr = gdr_backward_ref(q[:, a:b], k[:, a:b], v[:, a:b], g[:, a:b], beta[:, a:b], do[:, a:b], dsT[n:n + 1], None, S0[n:n + 1])
        for r, x in zip(refs[:5], gr[:5]):
            r[:, a:b] = x
        refs[5][n] = gr[5][0]
    for name, x, r in zip(("dq", "dk", "dv", "dg", "dbeta", "dS0"), (qd, kd, vd, gd, bd, sd), refs):
        assert max_rel_err(x.grad.float().cpu(), r.float()) <= TOL, (name, case)


@pytest.mark.parametrize("case", _shapes(10, 4242))
def test_random_shapes_backward(op, case):
    B, T, H, V, C, corr, _, nseg, seed = case
    q, k, v, g, beta, S0 = make_inputs(B, T, H, 64, V, seed=seed, frame_tokens=C, correlated=corr, dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(seed + 1)
    do = torch.randn(B, T, H, V, generator=gen).bfloat16()
    dsT = torch.randn(B, H, 64, V, generator=gen)
    ref = gdr_backward_ref(q, k, v, g, beta, do, dsT, None, S0)
    qd, kd, vd, gd, bd, sd = (x.cuda() for x in (q, k, v, g, beta, S0))
    _, _, cs = torch.ops.gdkvm.gdr_lkva_train(qd, kd, vd, gd, bd, None, sd, 0)
    got = torch.ops.gdkvm.gdr_lkva_bwd(qd, kd, vd, gd, bd, cs, do.cuda(), dsT.cuda(), 0.125, True, None, nseg << 8)
    torch.cuda.synchronize()
    for name, a, r in zip(("dq", "dk", "dv", "dg", "dbeta", "dS0"), got, ref):
        assert max_rel_err(a.float().cpu(), r.float()) <= TOL, (name, case)
