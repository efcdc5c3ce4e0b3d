"""ctypes view of include/gdkvm_gdr.h.  The library is loaded lazily and loudly: there is no
fallback implementation behind it."""
from __future__ import annotations

import ctypes
import os
import threading

from ._build import LIB_PATH

GDKVM_ABI_VERSION = 1
GDKVM_F32, GDKVM_BF16 = 0, 1
FLAG_FORCE_RECURRENT, FLAG_FORCE_CHUNKED, FLAG_FLAT_CHUNKS, FLAG_FRAME_CHUNKS = 0x1, 0x2, 0x4, 0x8


def FLAG_SEGMENTS(n: int) -> int:
    """GDKVM_FLAG_SEGMENTS(n): cut every chain into n (1..15) time segments; 0 = the library chooses."""
    return (int(n) & 0xF) << 8

EXPORTED_SYMBOLS = (
    "gdkvm_abi_version", "gdkvm_strerror", "gdkvm_last_cuda_error",
    "gdkvm_gdr_fwd", "gdkvm_gdr_fwd_varlen", "gdkvm_gdr_plan", "gdkvm_gdr_plan_segments", "gdkvm_gdr_plan_units", "gdkvm_gdr_plan_reason",
    "gdkvm_launch_count", "gdkvm_l2norm_fwd", "gdkvm_gdr_fwd_train", "gdkvm_gdr_chunk_states_bytes", "gdkvm_gdr_bwd",
    "gdkvm_gdr_fwd_train_varlen", "gdkvm_gdr_chunk_states_bytes_varlen",
    "gdkvm_qkvgb_project_fwd",
)


class GdkvmGdrParams(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("flags", ctypes.c_uint32),
        ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p),
        ("g", ctypes.c_void_p), ("beta", ctypes.c_void_p),
        ("initial_state", ctypes.c_void_p), ("o", ctypes.c_void_p), ("final_state", ctypes.c_void_p),
        ("q_stride", ctypes.c_int64 * 3), ("k_stride", ctypes.c_int64 * 3), ("v_stride", ctypes.c_int64 * 3),
        ("o_stride", ctypes.c_int64 * 3), ("g_stride", ctypes.c_int64 * 3), ("beta_stride", ctypes.c_int64 * 3),
        ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("H", ctypes.c_int32),
        ("K", ctypes.c_int32), ("V", ctypes.c_int32),
        ("frame_tokens", ctypes.c_int32), ("io_dtype", ctypes.c_int32), ("gate_dtype", ctypes.c_int32),
        ("scale", ctypes.c_float), ("reserved", ctypes.c_int32),
    ]


class GdkvmGdrBwdParams(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("flags", ctypes.c_uint32),
        ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p), ("g", ctypes.c_void_p), ("beta", ctypes.c_void_p),
        ("d_o", ctypes.c_void_p), ("d_final_state", ctypes.c_void_p), ("chunk_states", ctypes.c_void_p),
        ("dq", ctypes.c_void_p), ("dk", ctypes.c_void_p), ("dv", ctypes.c_void_p), ("dg", ctypes.c_void_p), ("dbeta", ctypes.c_void_p),
        ("d_initial_state", ctypes.c_void_p),
        ("q_stride", ctypes.c_int64 * 3), ("k_stride", ctypes.c_int64 * 3), ("v_stride", ctypes.c_int64 * 3),
        ("do_stride", ctypes.c_int64 * 3), ("g_stride", ctypes.c_int64 * 3), ("beta_stride", ctypes.c_int64 * 3),
        ("dq_stride", ctypes.c_int64 * 3), ("dk_stride", ctypes.c_int64 * 3), ("dv_stride", ctypes.c_int64 * 3),
        ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("H", ctypes.c_int32), ("K", ctypes.c_int32), ("V", ctypes.c_int32),
        ("io_dtype", ctypes.c_int32), ("gate_dtype", ctypes.c_int32), ("scale", ctypes.c_float),
        ("cu_seqlens", ctypes.c_void_p), ("cu_seqlens_bytes", ctypes.c_int32), ("n_seqs", ctypes.c_int32),
    ]


class GdkvmProjParams(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("flags", ctypes.c_uint32),
        ("x", ctypes.c_void_p), ("w", ctypes.c_void_p), ("bias", ctypes.c_void_p),
        ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p), ("g", ctypes.c_void_p), ("beta", ctypes.c_void_p),
        ("R", ctypes.c_int64), ("x_row_stride", ctypes.c_int64),
        ("D", ctypes.c_int32), ("H", ctypes.c_int32), ("K", ctypes.c_int32), ("V", ctypes.c_int32),
        ("eps", ctypes.c_float), ("reserved", ctypes.c_int32),
    ]


_lib = None
_lock = threading.Lock()


def load() -> ctypes.CDLL:
    """dlopen gdkvm_b200/libgdkvm_gdr.so; raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                    "gdkvm_b200 has no CPU or PyTorch fallback for the GDR/LKVA op.")
            lib = ctypes.CDLL(LIB_PATH)
            lib.gdkvm_abi_version.restype = ctypes.c_int
            lib.gdkvm_strerror.restype = ctypes.c_char_p
            lib.gdkvm_strerror.argtypes = [ctypes.c_int]
            lib.gdkvm_last_cuda_error.restype = ctypes.c_int
            lib.gdkvm_gdr_fwd.restype = ctypes.c_int
            lib.gdkvm_gdr_fwd.argtypes = [ctypes.POINTER(GdkvmGdrParams), ctypes.c_void_p]
            lib.gdkvm_gdr_fwd_varlen.restype = ctypes.c_int
            lib.gdkvm_gdr_fwd_varlen.argtypes = [ctypes.POINTER(GdkvmGdrParams), ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                                 ctypes.c_void_p]
            lib.gdkvm_gdr_plan.restype = ctypes.c_int
            lib.gdkvm_gdr_plan.argtypes = [ctypes.POINTER(GdkvmGdrParams)]
            lib.gdkvm_gdr_plan_segments.restype = ctypes.c_int
            lib.gdkvm_gdr_plan_segments.argtypes = [ctypes.POINTER(GdkvmGdrParams), ctypes.c_int]
            lib.gdkvm_gdr_plan_units.restype = ctypes.c_int
            lib.gdkvm_gdr_plan_units.argtypes = [ctypes.POINTER(GdkvmGdrParams), ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
            lib.gdkvm_gdr_plan_reason.restype = ctypes.c_char_p
            lib.gdkvm_gdr_plan_reason.argtypes = [ctypes.POINTER(GdkvmGdrParams)]
            lib.gdkvm_launch_count.restype = ctypes.c_uint64
            lib.gdkvm_l2norm_fwd.restype = ctypes.c_int
            lib.gdkvm_l2norm_fwd.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64,
                                             ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p]
            lib.gdkvm_gdr_fwd_train.restype = ctypes.c_int
            lib.gdkvm_gdr_fwd_train.argtypes = [ctypes.POINTER(GdkvmGdrParams), ctypes.c_void_p, ctypes.c_void_p]
            lib.gdkvm_gdr_chunk_states_bytes.restype = ctypes.c_int64
            lib.gdkvm_gdr_chunk_states_bytes.argtypes = [ctypes.c_int32] * 5
            lib.gdkvm_gdr_fwd_train_varlen.restype = ctypes.c_int
            lib.gdkvm_gdr_fwd_train_varlen.argtypes = [ctypes.POINTER(GdkvmGdrParams), ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                                       ctypes.c_void_p, ctypes.c_void_p]
            lib.gdkvm_gdr_chunk_states_bytes_varlen.restype = ctypes.c_int64
            lib.gdkvm_gdr_chunk_states_bytes_varlen.argtypes = [ctypes.c_int32] * 5
            lib.gdkvm_gdr_bwd.restype = ctypes.c_int
            lib.gdkvm_gdr_bwd.argtypes = [ctypes.POINTER(GdkvmGdrBwdParams), ctypes.c_void_p]
            lib.gdkvm_qkvgb_project_fwd.restype = ctypes.c_int
            lib.gdkvm_qkvgb_project_fwd.argtypes = [ctypes.POINTER(GdkvmProjParams), ctypes.c_void_p]
            if lib.gdkvm_abi_version() != GDKVM_ABI_VERSION:
                raise RuntimeError("libgdkvm_gdr.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


def strerror(rc: int) -> str:
    return load().gdkvm_strerror(rc).decode()


def launch_count() -> int:
    return int(load().gdkvm_launch_count())
