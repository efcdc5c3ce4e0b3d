"""torch.library.opcheck on every registered op (GPU): the schema says the truth about mutation / aliasing, the fake (meta)
implementation returns what the CUDA implementation returns (shapes, dtypes, strides, device), and the autograd registration is
well formed -- what a caller needs to put the ops under torch.compile / FakeTensor tracing of the SURROUNDING model (the ops
themselves stay opaque calls into the C ABI)."""
import pytest
import torch

from oracle.gdr_ref import make_inputs

pytestmark = pytest.mark.gpu
CHECKS = ("test_schema", "test_faketensor", "test_autograd_registration")


@pytest.fixture(scope="module")
def op(built_lib):
    assert torch.cuda.is_available(), "-m gpu tests need a B200"
    import gdkvm_b200
    return gdkvm_b200


def _inputs(B=2, T=130, H=2, V=128, seed=5):
    return [t.cuda() for t in make_inputs(B, T, H, 64, V, seed=seed, dtype=torch.bfloat16)]


def test_opcheck_forward_ops(op):
    q, k, v, g, beta, S0 = _inputs()
    torch.library.opcheck(torch.ops.gdkvm.gdr_lkva, (q, k, v, g, beta, None, S0, True, 0, 0), test_utils=CHECKS)
    torch.library.opcheck(torch.ops.gdkvm.gdr_lkva, (q, k, v, g, beta, 0.1, None, False, 65, 0), test_utils=CHECKS)
    cu = torch.tensor([0, 70, 70, 130], dtype=torch.int32, device="cuda")
    pk = lambda t: t[:1].contiguous()
    S3 = torch.randn(3, 2, 64, 128, device="cuda")
    torch.library.opcheck(torch.ops.gdkvm.gdr_lkva_varlen, (pk(q), pk(k), pk(v), pk(g), pk(beta), cu, None, S3, True, 0), test_utils=CHECKS)
    torch.library.opcheck(torch.ops.gdkvm.l2norm, (q,), test_utils=CHECKS)
    torch.library.opcheck(torch.ops.gdkvm.l2norm, (torch.randn(7, 3, 128, device="cuda"), 1e-5), test_utils=CHECKS)


def test_opcheck_training_ops(op):
    q, k, v, g, beta, S0 = _inputs()
    torch.library.opcheck(torch.ops.gdkvm.gdr_lkva_train, (q, k, v, g, beta, None, S0, 0), test_utils=CHECKS)
    o, sT, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
    do, dsT = torch.randn_like(o), torch.randn_like(sT)
    torch.library.opcheck(torch.ops.gdkvm.gdr_lkva_bwd, (q, k, v, g, beta, cs, do, dsT, 0.125, True, None, 0), test_utils=CHECKS)
    torch.library.opcheck(torch.ops.gdkvm.gdr_lkva_bwd, (q, k, v, g, beta, cs, do, None, 0.125, False, None, 0), test_utils=CHECKS)
    cu = torch.tensor([0, 64, 130], dtype=torch.int64, device="cuda")
    pk = lambda t: t[:1].contiguous()
    S2 = torch.randn(2, 2, 64, 128, device="cuda")
    torch.library.opcheck(torch.ops.gdkvm.gdr_lkva_varlen_train, (pk(q), pk(k), pk(v), pk(g), pk(beta), cu, None, S2, 0), test_utils=CHECKS)


def test_opcheck_projection(op):
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, 70, 128, generator=gen, device="cuda").bfloat16()
    w = (torch.randn(4 * (128 + 64) + 8, 128, generator=gen, device="cuda") / 11).bfloat16()
    b = torch.randn(w.shape[0], generator=gen, device="cuda")
    torch.library.opcheck(torch.ops.gdkvm.qkvgb_project, (x, w, b, 4, 64, 64), test_utils=CHECKS)
    torch.library.opcheck(torch.ops.gdkvm.qkvgb_project, (x, w, None, 4, 64, 64, 1e-5), test_utils=CHECKS)


def test_surrounding_model_traces_with_fake_tensors(op):
    """The skeleton's forward under FakeTensorMode on the GPU device type: every op answers through its fake implementation
    (no kernel runs), shapes and dtypes equal the real run's."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from gdkvm_b200.model import GDKVMSkeleton
    torch.manual_seed(0)
    model = GDKVMSkeleton().cuda().to(torch.bfloat16).eval()
    clip = torch.randn(1, 4, 1, 112, 112, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        logits, state = model(clip)
    n0 = op.launch_count()
    with FakeTensorMode(allow_non_fake_inputs=True) as mode, torch.no_grad():
        fl, fs = model(mode.from_tensor(clip))
    assert op.launch_count() == n0
    assert fl.shape == logits.shape and fl.dtype == logits.dtype and fs.shape == state.shape and fs.dtype == state.dtype
