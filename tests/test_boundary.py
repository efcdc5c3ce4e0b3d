"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, its struct matches the ctypes mirror, and validation/dispatch behave (no GPU compute)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gdkvm_gdr.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gdkvm_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    from gdkvm_b200 import _cabi
    lib = ctypes.CDLL(built_lib)
    names = _declared_functions()
    assert set(names) == set(_cabi.EXPORTED_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gdkvm_gdr.h but not exported"
    assert lib.gdkvm_abi_version() == _cabi.GDKVM_ABI_VERSION


@pytest.mark.parametrize("struct", ["GdkvmGdrParams", "GdkvmGdrBwdParams", "GdkvmProjParams"])
def test_struct_layout_matches_header(built_lib, tmp_path, struct):
    """Compile a probe against the real header with gcc and compare size/offsets with ctypes."""
    from gdkvm_b200 import _cabi
    cls = getattr(_cabi, struct)
    fields = [f[0] for f in cls._fields_]
    body = "".join(f'printf("{f} %zu\\n", offsetof({struct}, {f}));' for f in fields)
    c = tmp_path / "probe.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gdkvm_gdr.h"\n'
                 'int main(void){printf("sizeof %zu\\n", sizeof(' + struct + '));' + body + 'return 0;}')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(cls)
    for f in fields:
        assert int(out[f]) == getattr(cls, f).offset, f


def test_training_and_projection_entry_points_validate_without_gpu(built_lib):
    from gdkvm_b200 import _cabi
    lib = _cabi.load()
    assert lib.gdkvm_gdr_chunk_states_bytes(64, 6272, 8, 64, 256) == 64 * 8 * 98 * 256 * 64 * 2
    assert lib.gdkvm_gdr_chunk_states_bytes(1, 65, 1, 64, 128) == 2 * 128 * 64 * 2 and lib.gdkvm_gdr_chunk_states_bytes(0, 1, 1, 64, 128) == 0
    bp = _cabi.GdkvmGdrBwdParams()
    assert lib.gdkvm_gdr_bwd(None, None) == -1 and lib.gdkvm_gdr_bwd(ctypes.byref(bp), None) == -2       # NULL, ABI guard
    bp.struct_size = ctypes.sizeof(bp)
    bp.B = bp.T = bp.H = 1; bp.K = 64; bp.V = 128; bp.io_dtype = 1
    assert lib.gdkvm_gdr_bwd(ctypes.byref(bp), None) == -1                                              # tensors missing
    pp = _cabi.GdkvmProjParams()
    assert lib.gdkvm_qkvgb_project_fwd(ctypes.byref(pp), None) == -2
    pp.struct_size = ctypes.sizeof(pp)
    pp.R, pp.D, pp.H, pp.K, pp.V = 10, 256, 8, 64, 256
    assert lib.gdkvm_qkvgb_project_fwd(ctypes.byref(pp), None) == -1
    buf = torch.zeros(1 << 16, dtype=torch.bfloat16)
    for n in ("x", "w", "q", "k", "v", "g", "beta"):
        setattr(pp, n, buf.data_ptr())
    pp.x_row_stride = 256
    pp.H = 3
    assert lib.gdkvm_qkvgb_project_fwd(ctypes.byref(pp), None) == -8                                    # odd head count: unsupported
    pp.H = 8
    if not torch.cuda.is_available():
        assert lib.gdkvm_qkvgb_project_fwd(ctypes.byref(pp), None) in (-6, -7)                           # no device: never a host path


def _inputs(B=2, T=98, H=2, K=64, V=256, dtype=torch.bfloat16):
    from oracle.gdr_ref import make_inputs
    return make_inputs(B, T, H, K, V, dtype=dtype)


def test_plan_and_validation(built_lib):
    import gdkvm_b200
    from gdkvm_b200 import _cabi
    from gdkvm_b200.ops import _make_params
    q, k, v, g, beta, _ = _inputs()
    assert gdkvm_b200.plan(q, k, v, g, beta, flags=_cabi.FLAG_FORCE_RECURRENT) == 0
    assert gdkvm_b200.plan(q, k, v, g, beta, frame_tokens=49) in (0, 1)
    lib = _cabi.load()

    def rc_of(**over):
        p = _make_params(q, k, v, g, beta, v, None, None, 1.0, over.pop("frame_tokens", 0), over.pop("flags", 0))
        for kk, vv in over.items():
            setattr(p, kk, vv)
        return lib.gdkvm_gdr_plan(ctypes.byref(p))

    assert rc_of(K=48) == -3                      # GDKVM_ERR_SHAPE
    assert rc_of(frame_tokens=50) == -3           # T % frame_tokens != 0
    assert rc_of(struct_size=8) == -2             # GDKVM_ERR_ABI
    assert rc_of(io_dtype=7) == -4                # GDKVM_ERR_DTYPE
    assert rc_of(q=q.data_ptr() + 2) == -5        # GDKVM_ERR_ALIGN
    assert rc_of(q=None) == -1                    # GDKVM_ERR_NULL
    assert rc_of(flags=3) == -8                   # both force flags
    assert lib.gdkvm_gdr_plan(None) == -1
    for code in range(0, -9, -1):
        assert len(_cabi.strerror(code)) > 1
    assert "unknown" in _cabi.strerror(-99)


def test_time_segment_planning(built_lib):
    """Host-side schedule simulation behind gdkvm_gdr_plan_segments: chains are cut in time only when the last wave of
    (clip, head) chains over the SMs is ragged, never into segments shorter than 8 chunks, and a forced count wins."""
    import gdkvm_b200 as G

    def segs(B, F, C, H, V=256, flags=0, sms=0):
        T = F * C
        q = torch.empty(B, T, H, 64, dtype=torch.bfloat16)
        v = torch.empty(B, T, H, V, dtype=torch.bfloat16)
        g = torch.empty(B, T, H)
        return G.plan_segments(q, q, v, g, g, frame_tokens=C, flags=flags, sm_count=sms)

    assert segs(64, 128, 49, 8) == 2            # configs[1]: 512 chains on 148 SMs = 3.46 waves -> 1024 units, 6.92 waves
    assert segs(32, 20, 1024, 8) == 4           # configs[2] shape: 256 chains = 1.73 waves
    assert segs(8, 128, 49, 8) == 1             # 64 chains < 148 SMs: nothing to balance
    assert segs(37, 100, 64, 4) == 1            # exactly one wave
    assert segs(64, 128, 49, 8, sms=128) == 1   # 512 chains on 128 SMs: four full waves
    assert segs(3, 9, 64, 2) == 1               # short chains stay whole (< 8 chunks per segment)
    assert segs(64, 128, 49, 8, flags=G.FLAG_SEGMENTS(5)) == 5
    assert segs(1, 3, 64, 1, flags=G.FLAG_SEGMENTS(15)) == 3      # at most one segment per chunk
    assert segs(2, 10, 64, 1, flags=G.FLAG_SEGMENTS(4)) == 4      # 10 chunks in 3+3+3+1
    assert segs(2, 10, 64, 1, flags=G.FLAG_SEGMENTS(6)) == 5      # 2 chunks per segment -> 5 segments, none empty
    assert segs(2, 40, 49, 2, V=40, flags=1) == 1                 # recurrent path: no segments


def test_varlen_validation_without_gpu(built_lib):
    """gdkvm_gdr_fwd_varlen checks its arguments before it touches the device: packed layout (B = 1), offset width,
    offset pointer; with everything valid and no sm_100 device it fails loudly (no CPU path)."""
    from gdkvm_b200 import _cabi, ops
    lib = _cabi.load()
    T, H = 130, 2
    q = torch.zeros(1, T, H, 64, dtype=torch.bfloat16)
    v = torch.zeros(1, T, H, 256, dtype=torch.bfloat16)
    g = torch.zeros(1, T, H)
    cu = torch.tensor([0, 100, 130], dtype=torch.int64)
    call = lambda p, ptr, nbytes, n: lib.gdkvm_gdr_fwd_varlen(ctypes.byref(p), ctypes.c_void_p(ptr), nbytes, n, None)
    p = ops._make_params(q, q, v, g, g, v, None, None, 0.125, 0, 0)
    assert call(p, 0, 8, 2) == -1                              # NULL offsets
    assert call(p, cu.data_ptr(), 2, 2) == -3                  # offsets must be 4 or 8 bytes wide
    assert call(p, cu.data_ptr(), 8, 0) == -3                  # at least one clip
    assert call(p, cu.data_ptr() + 4, 8, 2) == -5              # misaligned offsets
    q2 = torch.zeros(2, T, H, 64, dtype=torch.bfloat16)
    v2 = torch.zeros(2, T, H, 256, dtype=torch.bfloat16)
    g2 = torch.zeros(2, T, H)
    p2 = ops._make_params(q2, q2, v2, g2, g2, v2, None, None, 0.125, 0, 0)
    assert call(p2, cu.data_ptr(), 8, 2) == -3                 # packed clips come as one batch row
    if not torch.cuda.is_available():
        assert call(p, cu.data_ptr(), 8, 2) in (-6, -7)        # no sm_100 device: ARCH or CUDA error, never a fallback
        with pytest.raises((RuntimeError, NotImplementedError)):
            ops.gdr_lkva_varlen(q, q, v, g, g, cu)


def test_fwd_without_gpu_fails_loudly(built_lib):
    """No GPU here: the C entry point must return an error, never compute on the host."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gdkvm_b200 import _cabi
    from gdkvm_b200.ops import _make_params
    q, k, v, g, beta, _ = _inputs(B=1, T=16, H=1)
    o = torch.full_like(v, 7.0)
    p = _make_params(q, k, v, g, beta, o, None, None, 1.0, 0, 0)
    rc = _cabi.load().gdkvm_gdr_fwd(ctypes.byref(p), None)
    assert rc in (-6, -7)                         # not sm_100 / CUDA error
    assert torch.all(o == 7.0)                    # untouched


def test_cpu_tensors_are_rejected(built_lib):
    import gdkvm_b200
    q, k, v, g, beta, _ = _inputs(B=1, T=16, H=1)
    with pytest.raises(NotImplementedError):
        gdkvm_b200.gdr_lkva(q, k, v, g, beta)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        gdkvm_b200.gdr_lkva_out(q, k, v, g, beta, torch.empty_like(v))


def test_fake_impl_shapes(built_lib):
    import gdkvm_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        q = torch.empty(3, 98, 4, 64, dtype=torch.bfloat16)
        v = torch.empty(3, 98, 4, 256, dtype=torch.bfloat16)
        g = torch.empty(3, 98, 4)
        o, sT = torch.ops.gdkvm.gdr_lkva(q, q, v, g, g)
        assert o.shape == (3, 98, 4, 256) and o.dtype == torch.bfloat16
        assert sT.shape == (3, 4, 64, 256) and sT.dtype == torch.float32


def test_argument_checks(built_lib):
    import gdkvm_b200
    q, k, v, g, beta, S0 = _inputs(B=1, T=16, H=1)
    with pytest.raises(ValueError):
        gdkvm_b200.plan(q, k[:, :8], v, g, beta)
    with pytest.raises(TypeError):
        gdkvm_b200.plan(q, k.float(), v, g, beta)
    with pytest.raises(NotImplementedError):
        gdkvm_b200.chunk_gated_delta_rule(q, k, v, g, beta, cu_seqlens=torch.tensor([0, 16]))


def test_product_does_not_import_oracle():
    """The shipped package must never route through the CPU oracle (or fla / triton)."""
    pkg = os.path.join(ROOT, "gdkvm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|fla|triton)\b", src, flags=re.M), f
    code = "import sys; import gdkvm_b200; assert not any(m.split('.')[0] in ('oracle','fla','triton') for m in sys.modules), 'leak'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
