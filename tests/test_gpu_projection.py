"""GPU tests of the fused projection prologue: the tcgen05 GEMM + normalising epilogue against
torch.nn.functional.linear in fp32 followed by the same formulas, and the projection -> memory op chain against the oracle."""
import pytest
import torch

from oracle.gdr_ref import gdr_recurrent_ref, max_rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def op(built_lib):
    assert torch.cuda.is_available(), "-m gpu tests need a B200"
    import gdkvm_b200
    return gdkvm_b200


def _fp32_reference(x, w, b, H, K, V, eps=1e-6):
    y = torch.nn.functional.linear(x.float(), w.float(), b)
    lead = x.shape[:-1]
    yq, yk, yv, yg, yb = torch.split(y, [H * K, H * K, H * V, H, H], dim=-1)
    nrm = lambda t: t.reshape(*lead, H, K) * torch.rsqrt(t.reshape(*lead, H, K).square().sum(-1, keepdim=True) + eps)
    return nrm(yq), nrm(yk), yv.reshape(*lead, H, V), torch.nn.functional.logsigmoid(yg), torch.sigmoid(yb)


def _make(lead, D, H, V, seed, bias):
    gen = torch.Generator().manual_seed(seed)
    N = H * (128 + V) + 2 * H
    x = torch.randn(*lead, D, generator=gen).bfloat16()
    w = (torch.randn(N, D, generator=gen) / D ** 0.5).bfloat16()
    b = 0.5 * torch.randn(N, generator=gen) if bias else None
    return x, w, b


@pytest.mark.parametrize("case", [
    # leading shape, D, H, V, bias
    ((300,), 256, 8, 256, True),          # ragged last row tile; the bench configuration
    ((3, 128), 64, 2, 64, False),         # one k-block, smallest head count
    ((2, 77), 512, 4, 128, True),         # eight k-blocks through the three-stage ring
    ((1, 5), 128, 6, 192, False),         # fewer rows than one tile; d_v = 192
    ((700,), 256, 8, 256, True),          # three 256-row blocks, the last ragged
])
@pytest.mark.parametrize("tile_rows", ["128", "256"])     # both tile shapes of the kernel (the library picks by problem size)
def test_projection_vs_fp32_linear(op, case, tile_rows, monkeypatch):
    lead, D, H, V, bias = case
    monkeypatch.setenv("GDKVM_PROJ_TILE_ROWS", tile_rows)
    if tile_rows == "256" and D > 256:
        with pytest.raises(RuntimeError, match="does not support"):
            x, w, b = _make(lead, D, H, V, 500 + D, bias)
            op.qkvgb_project(x.cuda(), w.cuda(), None, H, 64, V)
        return
    x, w, b = _make(lead, D, H, V, 500 + D, bias)
    ref = _fp32_reference(x, w, b, H, 64, V)
    got = op.qkvgb_project(x.cuda(), w.cuda(), b.cuda() if b is not None else None, H, 64, V)
    torch.cuda.synchronize()
    names = ("q", "k", "v", "g", "beta")
    for n, a, r in zip(names, got, ref):
        assert a.shape == r.shape, n
        err = (a.float().cpu() - r).abs().max().item()
        tol = {"q": 2.0 ** -8, "k": 2.0 ** -8, "v": 2.0 ** -8 * max(1.0, r.abs().max().item()), "g": 2e-4, "beta": 2e-4}[n]
        assert err <= tol, (n, err, tol)
    assert got[0].dtype == torch.bfloat16 and got[3].dtype == torch.float32
    # q, k rows have unit norm; the torch composition of the same map agrees
    assert (got[0].float().norm(dim=-1) - 1).abs().max().item() < 1e-2
    lib = op.qkvgb_project_reference(x.cuda(), w.cuda(), b.cuda() if b is not None else None, H, 64, V)
    for n, a, r in zip(names, got, lib):
        assert max_rel_err(a, r) <= 2e-2, n


def test_projection_feeds_the_memory_op(op):
    """features -> fused projection -> gdr_lkva, against the fp32 projection -> oracle recurrence."""
    B, T, D, H, V = 2, 3 * 49, 256, 2, 256
    x, w, b = _make((B, T), D, H, V, 600, True)
    q, k, v, g, beta = op.qkvgb_project(x.cuda(), w.cuda(), b.cuda(), H, 64, V)
    o, sT = op.gdr_lkva(q, k, v, g, beta, None, None, True, 49)
    qr, kr, vr, gr, br = _fp32_reference(x, w, b, H, 64, V)
    o_ref, s_ref = gdr_recurrent_ref(qr, kr, vr, gr, br, None, None)
    assert max_rel_err(o, o_ref) <= 2e-2 and max_rel_err(sT, s_ref) <= 2e-2


def test_projection_is_differentiable_and_rejects_bad_shapes(op):
    x, w, b = _make((40,), 128, 2, 128, 700, True)
    xd, wd, bd = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    outs = op.qkvgb_project(xd, wd, bd, 2, 64, 128)
    sum((o.float() * (i + 1)).sum() for i, o in enumerate(outs)).backward()
    x2, w2, b2 = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    outs2 = op.qkvgb_project_reference(x2, w2, b2, 2, 64, 128)
    sum((o.float() * (i + 1)).sum() for i, o in enumerate(outs2)).backward()
    for a, r in ((xd, x2), (wd, w2), (bd, b2)):
        assert max_rel_err(a.grad.float(), r.grad.float()) <= 1e-5
    with pytest.raises(ValueError):
        op.qkvgb_project(x.cuda(), w.cuda()[:-1], None, 2, 64, 128)
    with pytest.raises(RuntimeError, match="does not support"):
        op.qkvgb_project(x.cuda(), torch.zeros(3 * (128 + 128) + 6, 128, dtype=torch.bfloat16, device="cuda"), None, 3, 64, 128)   # odd H


def test_projection_full_size_properties(op):
    """configs[1] geometry (401 408 tokens, D = 256, 8 heads, d_v = 256 -> 3 088 columns; 3 136 row blocks over all SMs): every q / k row
    has unit norm, gates are <= 0 and beta in (0, 1), and three row blocks (first, one in the middle of a CTA's walk, the ragged
    last one) agree with the fp32 reference."""
    R, D, H, V = 64 * 6272 - 37, 256, 8, 256                     # a ragged last row block
    gen = torch.Generator(device="cuda").manual_seed(900)
    x = torch.randn(R, D, generator=gen, device="cuda").bfloat16()
    w = (torch.randn(H * (128 + V) + 2 * H, D, generator=gen, device="cuda") / D ** 0.5).bfloat16()
    b = 0.1 * torch.randn(w.shape[0], generator=gen, device="cuda")
    q, k, v, g, beta = op.qkvgb_project(x, w, b, H, 64, V)
    torch.cuda.synchronize()
    for t in (q, k):
        n = t.float().norm(dim=-1)
        assert float((n - 1).abs().max()) < 1e-2
    assert float(g.max()) <= 0 and float(beta.min()) > 0 and float(beta.max()) < 1 and bool(torch.isfinite(v.float()).all())
    for r0 in (0, 128 * 1777, (R // 128) * 128):
        sl = slice(r0, min(R, r0 + 128))
        ref = _fp32_reference(x[sl].cpu(), w.cpu(), b.cpu(), H, 64, V)
        for name, a, r in zip(("q", "k", "v", "g", "beta"), (q, k, v, g, beta), ref):
            tol = 2.0 ** -8 * max(1.0, float(r.abs().max())) if name in ("q", "k", "v") else 2e-4
            assert float((a[sl].float().cpu() - r).abs().max()) <= tol, (name, r0)


def test_projection_tile_shapes_agree_bit_for_bit(op, monkeypatch):
    """128-row and 256-row tiles run the same MMA sequence per output element (fp32 accumulation over the k-blocks in the same
    order) and the same epilogue arithmetic: identical bits, with and without a bias, ragged last block included."""
    x, w, b = _make((1500,), 256, 8, 256, 77, True)
    for bias in (None, b.cuda()):
        outs = {}
        for rows in ("128", "256"):
            monkeypatch.setenv("GDKVM_PROJ_TILE_ROWS", rows)
            outs[rows] = op.qkvgb_project(x.cuda(), w.cuda(), bias, 8, 64, 256)
        for name, a, c in zip(("q", "k", "v", "g", "beta"), outs["128"], outs["256"]):
            assert torch.equal(a, c), name
