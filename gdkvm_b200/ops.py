"""torch.library boundary of the GDKVM memory op.

``torch.ops.gdkvm.gdr_lkva`` keeps the call surface BASELINE.json ``north_star`` fixes for the
reference memory module -- ``(q, k, v, gate, beta, initial_state) -> (readout, final_state)`` --
which is argument-compatible with ``fla.ops.gated_delta_rule.chunk_gated_delta_rule``
(fla/ops/gated_delta_rule/chunk.py:365-377), so the upstream encoder -> KPFF -> memory -> decoder
pipeline can swap one import.  The reference tree itself has no code for this path
(reference README.md:1,36-38); the concept is named at README.md:20 and
website/src/content/homepage/en.json:20.

Only a CUDA implementation is registered.  There is deliberately no CPU kernel, no Triton and no
flash-linear-attention dispatch: CPU tensors raise from the dispatcher, and a missing
``libgdkvm_gdr.so`` raises from the loader.
"""
from __future__ import annotations

import ctypes
import math
import os
import warnings
from typing import Optional, Tuple

import torch

from . import _cabi

__all__ = ["gdr_lkva", "gdr_lkva_out", "train_unsupported_reason", "qkvgb_project", "qkvgb_project_reference", "gdr_lkva_varlen", "gdr_lkva_varlen_out", "check_inputs", "chunk_gated_delta_rule", "l2norm", "plan", "plan_reason", "plan_units",
           "plan_segments", "launch_count"]

_DT = {torch.float32: _cabi.GDKVM_F32, torch.bfloat16: _cabi.GDKVM_BF16}


def _strides3(t: torch.Tensor):
    s = t.stride()
    return (ctypes.c_int64 * 3)(s[0], s[1], s[2])


def _make_params(q, k, v, g, beta, o, s0, sT, scale, frame_tokens, flags) -> _cabi.GdkvmGdrParams:
    B, T, H, K = k.shape
    V = v.shape[-1]
    p = _cabi.GdkvmGdrParams()
    p.struct_size = ctypes.sizeof(_cabi.GdkvmGdrParams)
    p.flags = int(flags)
    p.q, p.k, p.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    p.g, p.beta = g.data_ptr(), beta.data_ptr()
    p.initial_state = s0.data_ptr() if s0 is not None else None
    p.o = o.data_ptr()
    p.final_state = sT.data_ptr() if sT is not None else None
    p.q_stride, p.k_stride, p.v_stride = _strides3(q), _strides3(k), _strides3(v)
    p.o_stride, p.g_stride, p.beta_stride = _strides3(o), _strides3(g), _strides3(beta)
    p.B, p.T, p.H, p.K, p.V = B, T, H, K, V
    p.frame_tokens = int(frame_tokens)
    p.io_dtype = _DT[q.dtype]
    p.gate_dtype = _DT[g.dtype]
    p.scale = float(scale)
    return p


def _same_device(q, **others):
    """Every tensor of a call must live on q's CUDA device: the op hands raw pointers to the kernel / to TMA, and a
    host pointer or another GPU's pointer there is an illegal-address fault (sticky: it kills the CUDA context)."""
    for name, t in others.items():
        if t is not None and t.device != q.device:
            raise ValueError(f"{name} is on {t.device} but q is on {q.device}: all tensors of one call must share one device")


_warned_fallback = set()
_CHECK_INPUTS = os.environ.get("GDKVM_CHECK_INPUTS", "0") not in ("", "0")


def _warn_fallback(lib, p, what):
    """Loud, once per reason: an un-forced call that the tcgen05 kernel cannot take runs the fp32 CUDA-core kernel."""
    if p.flags & (_cabi.FLAG_FORCE_RECURRENT | _cabi.FLAG_FORCE_CHUNKED) or p.T == 0:
        return
    if lib.gdkvm_gdr_plan(ctypes.byref(p)) != 0:         # the chunk kernel takes it, or the call is about to fail validation
        return
    reason = lib.gdkvm_gdr_plan_reason(ctypes.byref(p)).decode()
    if reason and reason not in _warned_fallback:
        _warned_fallback.add(reason)
        warnings.warn(f"{what}: running the token-recurrent fp32 CUDA-core kernel (several times slower than the tcgen05 "
                      f"chunk kernel) because {reason}", RuntimeWarning, stacklevel=3)


def _check(q, k, v, g, beta, initial_state):
    _same_device(q, k=k, v=v, g=g, beta=beta, initial_state=initial_state)
    if q.dim() != 4 or k.shape != q.shape or v.dim() != 4 or v.shape[:3] != q.shape[:3]:
        raise ValueError("expected q,k [B,T,H,K] and v [B,T,H,V]")
    if g.shape != q.shape[:3] or beta.shape != q.shape[:3]:
        raise ValueError("expected g,beta [B,T,H]")
    if q.dtype not in _DT or k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError("q,k,v must share one dtype of {float32, bfloat16}")
    if g.dtype not in _DT or beta.dtype != g.dtype:
        raise TypeError("g,beta must share one dtype of {float32, bfloat16}")
    for name, t in (("q", q), ("k", k), ("v", v)):
        if t.stride(-1) != 1:
            raise ValueError(f"{name}: innermost dimension must be contiguous")
    if initial_state is not None:
        B, T, H, K = q.shape
        if initial_state.shape != (B, H, K, v.shape[-1]) or initial_state.dtype != torch.float32:
            raise ValueError("initial_state must be fp32 [B,H,K,V]")


torch.library.define(
    "gdkvm::gdr_lkva",
    "(Tensor q, Tensor k, Tensor v, Tensor g, Tensor beta, float? scale=None, "
    "Tensor? initial_state=None, bool output_final_state=True, int frame_tokens=0, int flags=0) "
    "-> (Tensor, Tensor)",
)


def gdr_lkva_out(q, k, v, g, beta, o, final_state=None, scale=None, initial_state=None,
                 frame_tokens=0, flags=0) -> None:
    """Launch the op into caller-owned ``o`` (and ``final_state``) on the current stream.

    This is the C-ABI call with torch tensors as the buffer owners: no allocation, no sync.
    Used by the torch.library op below and by the host-buffer pipeline (gdkvm_b200.host).
    """
    _check(q, k, v, g, beta, initial_state)
    if not q.is_cuda:
        raise RuntimeError("gdkvm_b200 runs on a B200 only; there is no CPU implementation of gdr_lkva")
    lib = _cabi.load()
    _same_device(q, o=o, final_state=final_state)
    B, T, H, K = k.shape
    V = v.shape[-1]
    if o.shape != (B, T, H, V) or o.dtype != q.dtype or o.stride(-1) != 1:
        raise ValueError("o must be [B,T,H,V] in q.dtype with a contiguous last dimension")
    if final_state is not None and (final_state.shape != (B, H, K, V) or final_state.dtype != torch.float32
                                    or not final_state.is_contiguous()):
        raise ValueError("final_state must be contiguous fp32 [B,H,K,V]")
    if initial_state is not None and not initial_state.is_contiguous():
        raise ValueError("initial_state must be contiguous")
    if scale is None:
        scale = 1.0 / math.sqrt(K)
    p = _make_params(q, k, v, g, beta, o, initial_state, final_state, scale, frame_tokens, flags)
    if _CHECK_INPUTS:
        check_inputs(k, beta)
    _warn_fallback(lib, p, "gdkvm_b200.gdr_lkva")
    with torch.cuda.device(q.device):
        rc = lib.gdkvm_gdr_fwd(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        extra = f" (cudaError {lib.gdkvm_last_cuda_error()})" if rc == -7 else ""
        raise RuntimeError(f"gdkvm_gdr_fwd: {_cabi.strerror(rc)}{extra}")


@torch.library.impl("gdkvm::gdr_lkva", "CUDA")
def _gdr_lkva_cuda(q, k, v, g, beta, scale=None, initial_state=None, output_final_state=True,
                   frame_tokens=0, flags=0):
    B, T, H, K = k.shape
    V = v.shape[-1]
    if initial_state is not None:
        initial_state = initial_state.contiguous()
    o = torch.empty((B, T, H, V), dtype=q.dtype, device=q.device)
    sT = torch.empty((B, H, K, V) if output_final_state else (0,), dtype=torch.float32, device=q.device)
    gdr_lkva_out(q, k, v, g, beta, o, sT if output_final_state else None, scale, initial_state,
                 frame_tokens, flags)
    return o, sT


@torch.library.register_fake("gdkvm::gdr_lkva")
def _gdr_lkva_fake(q, k, v, g, beta, scale=None, initial_state=None, output_final_state=True,
                   frame_tokens=0, flags=0):
    B, T, H, K = k.shape
    V = v.shape[-1]
    o = q.new_empty((B, T, H, V))
    sT = q.new_empty((B, H, K, V) if output_final_state else (0,), dtype=torch.float32)
    return o, sT


# ---------------------------------------------------------------------------------------------------------------------
# training: forward that keeps the chunk-start states, backward kernel, autograd
# ---------------------------------------------------------------------------------------------------------------------
torch.library.define(
    "gdkvm::gdr_lkva_train",
    "(Tensor q, Tensor k, Tensor v, Tensor g, Tensor beta, float? scale=None, Tensor? initial_state=None, int flags=0) "
    "-> (Tensor, Tensor, Tensor)",
)
torch.library.define(
    "gdkvm::gdr_lkva_bwd",
    "(Tensor q, Tensor k, Tensor v, Tensor g, Tensor beta, Tensor chunk_states, Tensor d_o, Tensor? d_final_state, float scale, "
    "bool need_d_initial_state, Tensor? cu_seqlens=None, int flags=0) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)",
)
torch.library.define(
    "gdkvm::gdr_lkva_varlen_train",
    "(Tensor q, Tensor k, Tensor v, Tensor g, Tensor beta, Tensor cu_seqlens, float? scale=None, Tensor? initial_state=None, "
    "int flags=0) -> (Tensor, Tensor, Tensor)",
)


def train_unsupported_reason(q, k, v) -> str:
    """Why the training path (forward that keeps chunk states + backward kernel) cannot take these tensors ("" if it can)."""
    if q.dtype != torch.bfloat16:
        return "the backward pass takes bf16 q/k/v (fp32 I/O has no training path)"
    if k.shape[-1] != 64 or v.shape[-1] not in (64, 128, 256):
        return "the backward pass is built for d_k = 64 and d_v in {64, 128, 256}"
    return ""


@torch.library.impl("gdkvm::gdr_lkva_train", "CUDA")
def _gdr_lkva_train_cuda(q, k, v, g, beta, scale=None, initial_state=None, flags=0):
    _check(q, k, v, g, beta, initial_state)
    B, T, H, K = k.shape
    V = v.shape[-1]
    why = train_unsupported_reason(q, k, v)
    if why:
        raise NotImplementedError(f"gdkvm_b200 training forward: {why}")
    lib = _cabi.load()
    if initial_state is not None:
        initial_state = initial_state.contiguous()
    o = torch.empty((B, T, H, V), dtype=q.dtype, device=q.device)
    sT = torch.empty((B, H, K, V), dtype=torch.float32, device=q.device)
    nbytes = lib.gdkvm_gdr_chunk_states_bytes(B, T, H, K, V)
    cs = torch.empty((B * H, (T + 63) // 64, V, K), dtype=torch.bfloat16, device=q.device)
    assert cs.numel() * 2 == nbytes
    if T == 0:
        sT.copy_(initial_state) if initial_state is not None else sT.zero_()
        return o, sT, cs
    if scale is None:
        scale = 1.0 / math.sqrt(K)
    p = _make_params(q, k, v, g, beta, o, initial_state, sT, scale, 0, flags & ~0xF)
    with torch.cuda.device(q.device):
        rc = lib.gdkvm_gdr_fwd_train(ctypes.byref(p), ctypes.c_void_p(cs.data_ptr()),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        extra = f" (cudaError {lib.gdkvm_last_cuda_error()})" if rc == -7 else ""
        raise RuntimeError(f"gdkvm_gdr_fwd_train: {_cabi.strerror(rc)}{extra}")
    return o, sT, cs


@torch.library.register_fake("gdkvm::gdr_lkva_train")
def _gdr_lkva_train_fake(q, k, v, g, beta, scale=None, initial_state=None, flags=0):
    B, T, H, K = k.shape
    V = v.shape[-1]
    return (q.new_empty((B, T, H, V)), q.new_empty((B, H, K, V), dtype=torch.float32),
            q.new_empty((B * H, (T + 63) // 64, V, K), dtype=torch.bfloat16))


@torch.library.impl("gdkvm::gdr_lkva_bwd", "CUDA")
def _gdr_lkva_bwd_cuda(q, k, v, g, beta, chunk_states, d_o, d_final_state, scale, need_d_initial_state, cu_seqlens=None, flags=0):
    _check(q, k, v, g, beta, None)
    _same_device(q, chunk_states=chunk_states, d_o=d_o, d_final_state=d_final_state, cu_seqlens=cu_seqlens)
    B, T, H, K = k.shape
    V = v.shape[-1]
    NS = B if cu_seqlens is None else cu_seqlens.numel() - 1          # number of states: clips, or packed clips
    if d_o.shape != v.shape or d_o.dtype != q.dtype:
        raise ValueError("d_o must have the shape and dtype of the readout")
    if d_o.stride(-1) != 1:
        d_o = d_o.contiguous()
    if d_final_state is not None:
        if d_final_state.shape != (NS, H, K, V):
            raise ValueError("d_final_state must be [B,H,K,V] ([n_seqs,H,K,V] for packed clips)")
        d_final_state = d_final_state.to(torch.float32).contiguous()
    dq, dk, dv = torch.empty_like(q, memory_format=torch.contiguous_format), torch.empty_like(k, memory_format=torch.contiguous_format), \
        torch.empty_like(v, memory_format=torch.contiguous_format)
    dg = torch.empty((B, T, H), dtype=torch.float32, device=q.device)
    db = torch.empty((B, T, H), dtype=torch.float32, device=q.device)
    ds0 = torch.empty((NS, H, K, V) if need_d_initial_state else (0,), dtype=torch.float32, device=q.device)
    if T == 0:
        if need_d_initial_state:
            ds0.copy_(d_final_state) if d_final_state is not None else ds0.zero_()
        return dq, dk, dv, dg, db, ds0
    lib = _cabi.load()
    p = _cabi.GdkvmGdrBwdParams()
    p.struct_size = ctypes.sizeof(_cabi.GdkvmGdrBwdParams)
    p.flags = int(flags) & 0xF00
    p.q, p.k, p.v, p.g, p.beta = q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr()
    p.d_o = d_o.data_ptr()
    p.d_final_state = d_final_state.data_ptr() if d_final_state is not None else None
    p.chunk_states = chunk_states.data_ptr()
    p.dq, p.dk, p.dv, p.dg, p.dbeta = dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dg.data_ptr(), db.data_ptr()
    p.d_initial_state = ds0.data_ptr() if need_d_initial_state else None
    p.q_stride, p.k_stride, p.v_stride, p.do_stride = _strides3(q), _strides3(k), _strides3(v), _strides3(d_o)
    p.g_stride, p.beta_stride = _strides3(g), _strides3(beta)
    p.dq_stride, p.dk_stride, p.dv_stride = _strides3(dq), _strides3(dk), _strides3(dv)
    p.B, p.T, p.H, p.K, p.V = B, T, H, K, V
    p.io_dtype, p.gate_dtype, p.scale = _DT[q.dtype], _DT[g.dtype], float(scale)
    if cu_seqlens is not None:
        cu = cu_seqlens.contiguous()
        p.cu_seqlens, p.cu_seqlens_bytes, p.n_seqs = cu.data_ptr(), cu.element_size(), NS
    with torch.cuda.device(q.device):
        rc = lib.gdkvm_gdr_bwd(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        extra = f" (cudaError {lib.gdkvm_last_cuda_error()})" if rc == -7 else ""
        raise RuntimeError(f"gdkvm_gdr_bwd: {_cabi.strerror(rc)}{extra}")
    return dq, dk, dv, dg, db, ds0


@torch.library.register_fake("gdkvm::gdr_lkva_bwd")
def _gdr_lkva_bwd_fake(q, k, v, g, beta, chunk_states, d_o, d_final_state, scale, need_d_initial_state, cu_seqlens=None, flags=0):
    B, T, H, K = k.shape
    V = v.shape[-1]
    NS = B if cu_seqlens is None else cu_seqlens.shape[0] - 1
    f32 = dict(dtype=torch.float32)
    return (torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), q.new_empty((B, T, H), **f32), q.new_empty((B, T, H), **f32),
            q.new_empty((NS, H, K, V) if need_d_initial_state else (0,), **f32))


@torch.library.impl("gdkvm::gdr_lkva_varlen_train", "CUDA")
def _gdr_lkva_varlen_train_cuda(q, k, v, g, beta, cu_seqlens, scale=None, initial_state=None, flags=0):
    _check(q, k, v, g, beta, None)
    _same_device(q, cu_seqlens=cu_seqlens, initial_state=initial_state)
    B, T, H, K = k.shape
    V = v.shape[-1]
    why = train_unsupported_reason(q, k, v)
    if why:
        raise NotImplementedError(f"gdkvm_b200 training forward: {why}")
    if B != 1 or cu_seqlens.dim() != 1 or cu_seqlens.numel() < 2 or cu_seqlens.dtype not in (torch.int32, torch.int64):
        raise ValueError("packed clips: q,k,v [1, total_tokens, H, *] and cu_seqlens int32/int64 [n_seqs + 1]")
    N = cu_seqlens.numel() - 1
    cu = cu_seqlens.contiguous()
    if initial_state is not None:
        if initial_state.shape != (N, H, K, V) or initial_state.dtype != torch.float32:
            raise ValueError("initial_state must be fp32 [n_seqs,H,K,V]")
        initial_state = initial_state.contiguous()
    lib = _cabi.load()
    o = torch.empty((1, T, H, V), dtype=q.dtype, device=q.device)
    sT = torch.empty((N, H, K, V), dtype=torch.float32, device=q.device)
    cs = torch.empty((T // 64 + N + 1, H, V, K), dtype=torch.bfloat16, device=q.device)
    assert cs.numel() * 2 == lib.gdkvm_gdr_chunk_states_bytes_varlen(T, N, H, K, V)
    if T == 0:
        sT.copy_(initial_state) if initial_state is not None else sT.zero_()
        return o, sT, cs
    if scale is None:
        scale = 1.0 / math.sqrt(K)
    p = _make_params(q, k, v, g, beta, o, initial_state, sT, scale, 0, flags & ~0xF)
    with torch.cuda.device(q.device):
        rc = lib.gdkvm_gdr_fwd_train_varlen(ctypes.byref(p), ctypes.c_void_p(cu.data_ptr()), cu.element_size(), N,
                                            ctypes.c_void_p(cs.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        extra = f" (cudaError {lib.gdkvm_last_cuda_error()})" if rc == -7 else ""
        raise RuntimeError(f"gdkvm_gdr_fwd_train_varlen: {_cabi.strerror(rc)}{extra}")
    return o, sT, cs


@torch.library.register_fake("gdkvm::gdr_lkva_varlen_train")
def _gdr_lkva_varlen_train_fake(q, k, v, g, beta, cu_seqlens, scale=None, initial_state=None, flags=0):
    B, T, H, K = k.shape
    V = v.shape[-1]
    N = cu_seqlens.shape[0] - 1
    return (q.new_empty((1, T, H, V)), q.new_empty((N, H, K, V), dtype=torch.float32),
            q.new_empty((T // 64 + N + 1, H, V, K), dtype=torch.bfloat16))


def _vtrain_setup(ctx, inputs, output):
    q, k, v, g, beta, cu_seqlens, scale, initial_state, flags = inputs
    ctx.save_for_backward(q, k, v, g, beta, output[2], cu_seqlens)
    ctx.scale = scale if scale is not None else 1.0 / math.sqrt(k.shape[-1])
    ctx.has_s0 = initial_state is not None
    ctx.set_materialize_grads(False)


def _vtrain_backward(ctx, d_o, d_sT, _d_cs):
    q, k, v, g, beta, cs, cu = ctx.saved_tensors
    if d_o is None:
        d_o = torch.zeros_like(v)
    dq, dk, dv, dg, db, ds0 = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, d_o.to(q.dtype), d_sT, ctx.scale, ctx.has_s0, cu)
    return dq, dk, dv, dg.to(g.dtype), db.to(beta.dtype), None, None, (ds0 if ctx.has_s0 else None), None


def _train_setup(ctx, inputs, output):
    q, k, v, g, beta, scale, initial_state, flags = inputs
    ctx.save_for_backward(q, k, v, g, beta, output[2])
    ctx.scale = scale if scale is not None else 1.0 / math.sqrt(k.shape[-1])
    ctx.has_s0 = initial_state is not None
    ctx.set_materialize_grads(False)


def _train_backward(ctx, d_o, d_sT, _d_cs):
    q, k, v, g, beta, cs = ctx.saved_tensors
    if d_o is None:
        d_o = torch.zeros_like(v)
    dq, dk, dv, dg, db, ds0 = torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, d_o.to(q.dtype), d_sT, ctx.scale, ctx.has_s0)
    return dq, dk, dv, dg.to(g.dtype), db.to(beta.dtype), None, (ds0 if ctx.has_s0 else None), None


torch.library.register_autograd("gdkvm::gdr_lkva_train", _train_backward, setup_context=_train_setup)
torch.library.register_autograd("gdkvm::gdr_lkva_varlen_train", _vtrain_backward, setup_context=_vtrain_setup)


def _no_backward(name):
    def backward(ctx, *grads):
        raise NotImplementedError(
            f"torch.ops.gdkvm.{name} has no backward formula: differentiate through gdkvm_b200.gdr_lkva / GDRMemory / "
            "gdr_lkva_varlen / chunk_gated_delta_rule (they route a call that needs gradients to the training forward: bf16, "
            "d_k = 64, d_v in {64, 128, 256})")
    return backward


torch.library.register_autograd("gdkvm::gdr_lkva", _no_backward("gdr_lkva"))


torch.library.define(
    "gdkvm::gdr_lkva_varlen",
    "(Tensor q, Tensor k, Tensor v, Tensor g, Tensor beta, Tensor cu_seqlens, float? scale=None, "
    "Tensor? initial_state=None, bool output_final_state=True, int flags=0) -> (Tensor, Tensor)",
)


def gdr_lkva_varlen_out(q, k, v, g, beta, cu_seqlens, o, final_state=None, scale=None, initial_state=None, flags=0) -> None:
    """The packed variable-length call into caller-owned ``o`` [1,T,H,V] (and ``final_state`` [N,H,K,V]): the C-ABI call
    with torch tensors as buffer owners.  Rows of ``o`` that belong to no clip (before ``cu_seqlens[0]``, from
    ``cu_seqlens[-1]`` on) are not written."""
    _check(q, k, v, g, beta, None)
    if not q.is_cuda:
        raise RuntimeError("gdkvm_b200 runs on a B200 only; there is no CPU implementation of gdr_lkva_varlen")
    _same_device(q, cu_seqlens=cu_seqlens, initial_state=initial_state, o=o, final_state=final_state)
    B, T, H, K = k.shape
    V = v.shape[-1]
    if B != 1:
        raise ValueError("packed variable-length clips: q,k,v must be [1, total_tokens, H, *]")
    if cu_seqlens.dim() != 1 or cu_seqlens.numel() < 2 or cu_seqlens.dtype not in (torch.int32, torch.int64):
        raise ValueError("cu_seqlens must be a 1-D int32/int64 tensor of n_seqs + 1 offsets")
    N = cu_seqlens.numel() - 1
    cu = cu_seqlens.contiguous()
    if o.shape != (1, T, H, V) or o.dtype != q.dtype or o.stride(-1) != 1:
        raise ValueError("o must be [1,T,H,V] in q.dtype with a contiguous last dimension")
    for name, st in (("initial_state", initial_state), ("final_state", final_state)):
        if st is not None and (st.shape != (N, H, K, V) or st.dtype != torch.float32 or not st.is_contiguous()):
            raise ValueError(f"{name} must be contiguous fp32 [n_seqs,H,K,V]")
    if scale is None:
        scale = 1.0 / math.sqrt(K)
    p = _make_params(q, k, v, g, beta, o, initial_state, final_state, scale, 0, flags)
    lib = _cabi.load()
    _warn_fallback(lib, p, "gdkvm_b200.gdr_lkva_varlen")
    with torch.cuda.device(q.device):
        rc = lib.gdkvm_gdr_fwd_varlen(ctypes.byref(p), ctypes.c_void_p(cu.data_ptr()), cu.element_size(), N,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        extra = f" (cudaError {lib.gdkvm_last_cuda_error()})" if rc == -7 else ""
        raise RuntimeError(f"gdkvm_gdr_fwd_varlen: {_cabi.strerror(rc)}{extra}")


@torch.library.impl("gdkvm::gdr_lkva_varlen", "CUDA")
def _gdr_lkva_varlen_cuda(q, k, v, g, beta, cu_seqlens, scale=None, initial_state=None, output_final_state=True, flags=0):
    B, T, H, K = k.shape
    V = v.shape[-1]
    N = cu_seqlens.numel() - 1
    if initial_state is not None:
        initial_state = initial_state.contiguous()
    o = torch.empty((1, T, H, V), dtype=q.dtype, device=q.device)
    sT = torch.empty((N, H, K, V) if output_final_state else (0,), dtype=torch.float32, device=q.device)
    gdr_lkva_varlen_out(q, k, v, g, beta, cu_seqlens, o, sT if output_final_state else None, scale, initial_state, flags)
    return o, sT


@torch.library.register_fake("gdkvm::gdr_lkva_varlen")
def _gdr_lkva_varlen_fake(q, k, v, g, beta, cu_seqlens, scale=None, initial_state=None, output_final_state=True, flags=0):
    B, T, H, K = k.shape
    V = v.shape[-1]
    N = cu_seqlens.shape[0] - 1
    return q.new_empty((1, T, H, V)), q.new_empty((N, H, K, V) if output_final_state else (0,), dtype=torch.float32)


torch.library.register_autograd("gdkvm::gdr_lkva_varlen", _no_backward("gdr_lkva_varlen"))

torch.library.define("gdkvm::l2norm", "(Tensor x, float eps=1e-6) -> Tensor")


@torch.library.impl("gdkvm::l2norm", "CUDA")
def _l2norm_cuda(x, eps=1e-6):
    if x.dtype not in _DT:
        raise TypeError("l2norm: x must be float32 or bfloat16")
    D = x.shape[-1]
    xc = x if x.is_contiguous() else x.contiguous()
    y = torch.empty_like(xc)
    rows = xc.numel() // D if D else 0
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.gdkvm_l2norm_fwd(ctypes.c_void_p(xc.data_ptr()), ctypes.c_void_p(y.data_ptr()), rows, D, D, D, _DT[x.dtype],
                                  float(eps), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        extra = f" (cudaError {lib.gdkvm_last_cuda_error()})" if rc == -7 else ""
        raise RuntimeError(f"gdkvm_l2norm_fwd: {_cabi.strerror(rc)}{extra}")
    return y


@torch.library.register_fake("gdkvm::l2norm")
def _l2norm_fake(x, eps=1e-6):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def _l2norm_setup(ctx, inputs, output):
    x, eps = inputs
    ctx.save_for_backward(x)
    ctx.eps = eps


def _l2norm_backward(ctx, dy):
    # y = x r, r = rsqrt(sum x^2 + eps):  dx = r (dy - y (y . dy))   (elementwise torch ops; not a hot path)
    (x,) = ctx.saved_tensors
    xf, dyf = x.float(), dy.float()
    r = torch.rsqrt(xf.square().sum(-1, keepdim=True) + ctx.eps)
    y = xf * r
    return (r * (dyf - y * (y * dyf).sum(-1, keepdim=True))).to(x.dtype), None


torch.library.register_autograd("gdkvm::l2norm", _l2norm_backward, setup_context=_l2norm_setup)


torch.library.define("gdkvm::qkvgb_project",
                     "(Tensor x, Tensor weight, Tensor? bias, int heads, int d_k, int d_v, float eps=1e-6) -> (Tensor, Tensor, Tensor, Tensor, Tensor)")


@torch.library.impl("gdkvm::qkvgb_project", "CUDA")
def _qkvgb_project_cuda(x, weight, bias, heads, d_k, d_v, eps=1e-6):
    H, K, V = int(heads), int(d_k), int(d_v)
    N = H * (2 * K + V) + 2 * H
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise TypeError("qkvgb_project: features and weight must be bfloat16")
    if weight.dim() != 2 or weight.shape[0] != N or weight.shape[1] != x.shape[-1]:
        raise ValueError(f"qkvgb_project: weight must be [H (2 d_k + d_v) + 2 H = {N}, D = {x.shape[-1]}] (rows q | k | v | g | beta)")
    _same_device(x, weight=weight, bias=bias)
    D = x.shape[-1]
    lead = x.shape[:-1]
    x2 = x.reshape(-1, D)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    w = weight.contiguous()
    if bias is not None:
        if bias.shape != (N,):
            raise ValueError("qkvgb_project: bias must be [N]")
        bias = bias.to(torch.float32).contiguous()
    R = x2.shape[0]
    dev = x.device
    q = torch.empty(*lead, H, K, dtype=torch.bfloat16, device=dev)
    k = torch.empty(*lead, H, K, dtype=torch.bfloat16, device=dev)
    v = torch.empty(*lead, H, V, dtype=torch.bfloat16, device=dev)
    g = torch.empty(*lead, H, dtype=torch.float32, device=dev)
    beta = torch.empty(*lead, H, dtype=torch.float32, device=dev)
    p = _cabi.GdkvmProjParams()
    p.struct_size = ctypes.sizeof(_cabi.GdkvmProjParams)
    p.x, p.w, p.bias = x2.data_ptr(), w.data_ptr(), (bias.data_ptr() if bias is not None else None)
    p.q, p.k, p.v, p.g, p.beta = q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr()
    p.R, p.x_row_stride, p.D, p.H, p.K, p.V, p.eps = R, x2.stride(0), D, H, K, V, float(eps)
    rows = os.environ.get("GDKVM_PROJ_TILE_ROWS", "")                                      # tests / A-B runs: force a tile shape
    if rows not in ("", "128", "256"):
        raise ValueError(f"GDKVM_PROJ_TILE_ROWS must be 128 or 256 (got {rows!r})")
    p.flags = {"": 0, "128": 1, "256": 2}[rows]
    lib = _cabi.load()
    with torch.cuda.device(dev):
        rc = lib.gdkvm_qkvgb_project_fwd(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        extra = f" (cudaError {lib.gdkvm_last_cuda_error()})" if rc == -7 else ""
        raise RuntimeError(f"gdkvm_qkvgb_project_fwd: {_cabi.strerror(rc)}{extra}")
    return q, k, v, g, beta


@torch.library.register_fake("gdkvm::qkvgb_project")
def _qkvgb_project_fake(x, weight, bias, heads, d_k, d_v, eps=1e-6):
    lead = x.shape[:-1]
    f32 = dict(dtype=torch.float32)
    return (x.new_empty(*lead, heads, d_k), x.new_empty(*lead, heads, d_k), x.new_empty(*lead, heads, d_v),
            x.new_empty(*lead, heads, **f32), x.new_empty(*lead, heads, **f32))


def qkvgb_project_reference(x, weight, bias, heads, d_k, d_v, eps=1e-6):
    """The same map in plain torch ops (library GEMM + elementwise passes): the unfused route bench.py times the fused kernel
    against, and the formula autograd differentiates in the backward of ``qkvgb_project``."""
    H, K, V = heads, d_k, d_v
    y = torch.nn.functional.linear(x, weight, bias.to(x.dtype) if bias is not None else None).float()
    lead = x.shape[:-1]
    yq, yk, yv, yg, yb = torch.split(y, [H * K, H * K, H * V, H, H], dim=-1)
    nrm = lambda t: (t.reshape(*lead, H, K) * torch.rsqrt(t.reshape(*lead, H, K).square().sum(-1, keepdim=True) + eps))
    return (nrm(yq).to(x.dtype), nrm(yk).to(x.dtype), yv.reshape(*lead, H, V).to(x.dtype),
            torch.nn.functional.logsigmoid(yg), torch.sigmoid(yb))


def _proj_setup(ctx, inputs, output):
    x, weight, bias, heads, d_k, d_v, eps = inputs
    ctx.save_for_backward(x, weight, bias)
    ctx.cfg = (heads, d_k, d_v, eps)


def _proj_backward(ctx, dq, dk, dv, dg, db):
    # recompute through the torch formula (library GEMMs): training goes through cuBLAS here, inference through the fused kernel
    x, weight, bias = ctx.saved_tensors
    with torch.enable_grad():
        xs = x.detach().requires_grad_(True)
        ws = weight.detach().requires_grad_(True)
        bs = bias.detach().requires_grad_(True) if bias is not None else None
        outs = qkvgb_project_reference(xs, ws, bs, *ctx.cfg)
        pairs = [(o, d) for o, d in zip(outs, (dq, dk, dv, dg, db)) if d is not None]
        gr = torch.autograd.grad([o for o, _ in pairs], [xs, ws] + ([bs] if bs is not None else []), [d for _, d in pairs], allow_unused=True)
    return gr[0], gr[1], (gr[2] if bs is not None else None), None, None, None, None


torch.library.register_autograd("gdkvm::qkvgb_project", _proj_backward, setup_context=_proj_setup)


def qkvgb_project(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], heads: int, d_k: int = 64, d_v: int = 256,
                  eps: float = 1e-6):
    """Fused projection prologue of the memory op: ``x [..., D]`` (bf16) times ``weight [H (2 d_k + d_v) + 2 H, D]`` (rows ordered
    q | k | v | g | beta) in one tcgen05 GEMM whose epilogue L2-normalises q and k per head, applies logsigmoid / sigmoid to
    the gate / beta columns and writes ``(q, k [..., H, d_k], v [..., H, d_v]`` bf16, ``g, beta [..., H]`` fp32) -- exactly
    the operands of ``gdr_lkva`` (reference: KPFF, website/src/content/homepage/en.json:20)."""
    return torch.ops.gdkvm.qkvgb_project(x, weight, bias, heads, d_k, d_v, eps)


def l2norm(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Row-wise ``x * rsqrt(sum(x^2, -1) + eps)`` on the B200 (the q/k normalisation in front of the memory op;
    fla's ``use_qk_l2norm_in_kernel``, fla/ops/gated_delta_rule/chunk.py:374).  D in {32, 64, 128, 256}."""
    return torch.ops.gdkvm.l2norm(x, eps)


def gdr_lkva(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, g: torch.Tensor, beta: torch.Tensor,
             scale: Optional[float] = None, initial_state: Optional[torch.Tensor] = None,
             output_final_state: bool = True, frame_tokens: int = 0,
             flags: int = 0) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """LKVA readout + GDR state update over a batch of clips on the current B200.

    q,k [B,T,H,K]; v [B,T,H,V]; g (log-space gate) and beta [B,T,H]; initial_state fp32 [B,H,K,V].
    Returns ``(o [B,T,H,V] in q.dtype, final_state fp32 [B,H,K,V] or None)``.
    ``frame_tokens=C`` declares T = F*C with every frame one chunk (north_star).
    Differentiable (bf16, d_k = 64, d_v in {64, 128, 256}): when gradients are enabled and an input requires them, the call
    runs the training forward and autograd runs the hand-written backward kernel (csrc/gdr_bwd_sm100.cu).
    """
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (q, k, v, g, beta, initial_state)):
        # a call that will be differentiated: the training forward (keeps the chunk-start states for the backward kernel)
        why = train_unsupported_reason(q, k, v)
        if why:
            raise NotImplementedError(f"gdkvm_b200.gdr_lkva cannot be differentiated for these tensors: {why}")
        o, sT, _ = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, scale, initial_state, flags)
        return o, (sT if output_final_state else None)
    o, sT = torch.ops.gdkvm.gdr_lkva(q, k, v, g, beta, scale, initial_state, output_final_state,
                                     frame_tokens, flags)
    return o, (sT if output_final_state else None)


def gdr_lkva_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, g: torch.Tensor, beta: torch.Tensor,
                    cu_seqlens: torch.Tensor, scale: Optional[float] = None, initial_state: Optional[torch.Tensor] = None,
                    output_final_state: bool = True, flags: int = 0) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """The memory op over PACKED clips of different lengths (fla's ``cu_seqlens``, fla/ops/gated_delta_rule/chunk.py:375).

    q,k [1,T,H,K]; v [1,T,H,V]; g,beta [1,T,H]; ``cu_seqlens`` int32/int64 [N+1] ON THE DEVICE (clip n = rows
    cu_seqlens[n] .. cu_seqlens[n+1]-1); initial_state fp32 [N,H,K,V].  Returns ``(o [1,T,H,V], final_state [N,H,K,V] or
    None)``.  Nothing is read back to the host: the work-unit table is built on the device.
    """
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (q, k, v, g, beta, initial_state)):
        why = train_unsupported_reason(q, k, v)
        if why:
            raise NotImplementedError(f"gdkvm_b200.gdr_lkva_varlen cannot be differentiated for these tensors: {why}")
        o, sT, _ = torch.ops.gdkvm.gdr_lkva_varlen_train(q, k, v, g, beta, cu_seqlens, scale, initial_state, flags)
        return o, (sT if output_final_state else None)
    o, sT = torch.ops.gdkvm.gdr_lkva_varlen(q, k, v, g, beta, cu_seqlens, scale, initial_state, output_final_state, flags)
    return o, (sT if output_final_state else None)


def chunk_gated_delta_rule(q, k, v, g, beta, scale=None, initial_state=None, output_final_state=False,
                           **kwargs):
    """Name- and argument-compatible alias of fla's entry point (fla/ops/gated_delta_rule/chunk.py:365)."""
    unsupported = {kk: vv for kk, vv in kwargs.items()
                   if kk not in ("frame_tokens", "flags", "use_qk_l2norm_in_kernel", "cu_seqlens") and vv is not None
                   and vv is not False}
    if unsupported:
        raise NotImplementedError(f"gdkvm_b200.chunk_gated_delta_rule: unsupported arguments {sorted(unsupported)}")
    if kwargs.get("use_qk_l2norm_in_kernel"):      # one streaming CUDA pass each (not yet fused into the chunk kernel)
        q, k = l2norm(q), l2norm(k)
    if kwargs.get("cu_seqlens") is not None:
        return gdr_lkva_varlen(q, k, v, g, beta, kwargs["cu_seqlens"], scale, initial_state, output_final_state,
                               kwargs.get("flags", 0))
    return gdr_lkva(q, k, v, g, beta, scale, initial_state, output_final_state,
                    kwargs.get("frame_tokens", 0), kwargs.get("flags", 0))


def check_inputs(k: torch.Tensor, beta: torch.Tensor) -> float:
    """Debugging aid for the one numerical precondition of the op: ``beta_i |k_i|^2 <= 2`` for every token.  Beyond it the
    delta rule itself is unstable (the factor ``I - beta k k^T`` has an eigenvalue below -1, so the state grows
    geometrically in exact arithmetic too), and the tcgen05 kernel's fp16 triangular solve overflows where the fp32
    recurrence would merely explode.  L2-normalised keys (``l2norm`` / ``use_qk_l2norm_in_kernel``) with beta in (0, 1)
    are always inside.  Synchronises; returns the maximum of ``beta |k|^2`` and raises ``ValueError`` above 2.
    Called on every op call when the environment variable ``GDKVM_CHECK_INPUTS=1`` is set."""
    m = float((beta.float() * k.float().square().sum(-1)).max()) if k.numel() else 0.0
    if not m <= 2.0:
        raise ValueError(f"gdkvm_b200: max beta |k|^2 = {m:.3g} > 2: the gated delta rule is unstable for these inputs "
                         "(normalise k, or scale beta by 1/|k|^2)")
    return m


def plan(q, k, v, g, beta, *, frame_tokens: int = 0, flags: int = 0) -> int:
    """Which kernel the library would pick (0 recurrent, 1 tcgen05 chunked); needs no GPU."""
    _check(q, k, v, g, beta, None)
    o = v  # same geometry as the output
    p = _make_params(q, k, v, g, beta, o, None, None, 1.0, frame_tokens, flags)
    rc = _cabi.load().gdkvm_gdr_plan(ctypes.byref(p))
    if rc < 0:
        raise RuntimeError(f"gdkvm_gdr_plan: {_cabi.strerror(rc)}")
    return rc


def plan_reason(q, k, v, g, beta, *, frame_tokens: int = 0, flags: int = 0) -> str:
    """Why the library would NOT take the tcgen05 chunk kernel for these tensors ("" when it would); needs no GPU."""
    _check(q, k, v, g, beta, None)
    p = _make_params(q, k, v, g, beta, v, None, None, 1.0, frame_tokens, flags)
    return _cabi.load().gdkvm_gdr_plan_reason(ctypes.byref(p)).decode()


def plan_segments(q, k, v, g, beta, *, frame_tokens: int = 0, flags: int = 0, sm_count: int = 0) -> int:
    """Time segments per (clip, head) chain the chunked kernel would use on a device with ``sm_count`` SMs
    (0 = 148, a B200); host arithmetic only."""
    _check(q, k, v, g, beta, None)
    p = _make_params(q, k, v, g, beta, v, None, None, 1.0, frame_tokens, flags)
    rc = _cabi.load().gdkvm_gdr_plan_segments(ctypes.byref(p), int(sm_count))
    if rc < 0:
        raise RuntimeError(f"gdkvm_gdr_plan_segments: {_cabi.strerror(rc)}")
    return rc


def plan_units(q, k, v, g, beta, *, frame_tokens: int = 0, flags: int = 0, sm_count: int = 0) -> dict:
    """The work units an inference call would be scheduled as (``gdkvm_gdr_plan_units``): ``{"mixed": bool, "units": n, "uncut_clips":
    a, "cut_clips": b, "segments": s}``; host arithmetic only."""
    _check(q, k, v, g, beta, None)
    p = _make_params(q, k, v, g, beta, v, None, None, 1.0, frame_tokens, flags)
    out = (ctypes.c_int32 * 4)()
    rc = _cabi.load().gdkvm_gdr_plan_units(ctypes.byref(p), int(sm_count), out)
    if rc < 0:
        raise RuntimeError(f"gdkvm_gdr_plan_units: {_cabi.strerror(rc)}")
    return {"mixed": rc == 1, "units": out[0], "uncut_clips": out[1], "cut_clips": out[2], "segments": out[3]}


def launch_count() -> int:
    return _cabi.launch_count()
