"""configs[4] skeleton (encoder -> KPFF -> projection -> memory -> decoder): shapes on the meta device (CPU, through the ops'
fake implementations), and one small run on the GPU against the unfused projection route."""
import pytest
import torch


def test_skeleton_shapes_on_meta():
    from gdkvm_b200.model import GDKVMSkeleton
    with torch.device("meta"):
        m = GDKVMSkeleton(heads=4, d_v=128)
        logits, state = m(torch.empty(2, 3, 1, 112, 112))
    assert logits.shape == (2, 3, 1, 112, 112) and state.shape == (2, 4, 64, 128) and state.dtype == torch.float32


@pytest.mark.gpu
def test_skeleton_runs_and_the_fused_projection_changes_nothing_material(built_lib):
    from gdkvm_b200.model import GDKVMSkeleton
    torch.manual_seed(0)
    m = GDKVMSkeleton().cuda().to(torch.bfloat16).eval()
    clip = torch.randn(2, 6, 1, 112, 112, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        a, sa = m(clip)
        first, s1 = m(clip[:, :3])
        second, s2 = m(clip[:, 3:], s1)                       # streaming: the state carries the memory across calls
        m.fused_projection = False
        b, sb = m(clip)
    assert a.shape == (2, 6, 1, 112, 112) and bool(torch.isfinite(a.float()).all())
    den = a.float().abs().max().clamp_min(1e-6)
    assert float((torch.cat([first, second], 1).float() - a.float()).abs().max() / den) < 5e-2
    assert float((a.float() - b.float()).abs().max() / den) < 5e-2
    assert float((sa - sb).abs().max() / sb.abs().max()) < 5e-2 and float((s2 - sa).abs().max() / sa.abs().max()) < 5e-2


@pytest.mark.gpu
def test_skeleton_trains_through_the_memory_kernels(built_lib):
    """One training step of the stand-in model: the loss reaches every parameter through the backward kernel of the memory op
    (and the recomputed backward of the fused projection); gradients are finite and not all zero."""
    from gdkvm_b200.model import GDKVMSkeleton
    torch.manual_seed(1)
    m = GDKVMSkeleton(heads=4, d_v=128).cuda().to(torch.bfloat16).train()
    clip = torch.randn(2, 4, 1, 112, 112, device="cuda", dtype=torch.bfloat16)
    target = (torch.rand(2, 4, 1, 112, 112, device="cuda") > 0.5).to(torch.bfloat16)
    logits, state = m(clip)
    assert logits.requires_grad and state.requires_grad
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.float(), target.float())
    loss.backward()
    torch.cuda.synchronize()
    for name, p_ in m.named_parameters():
        assert p_.grad is not None and bool(torch.isfinite(p_.grad.float()).all()), name
    assert float(m.proj_weight.grad.float().abs().max()) > 0 and float(m.encoder.stem[0][0].weight.grad.float().abs().max()) > 0
