"""ctypes binding of oracle/gdr_ref.c (TEST INFRASTRUCTURE ONLY -- see that file's header)."""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgdr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile gdr_ref.c -> libgdr_oracle.so (gcc, pthreads).  Building the checker is not using it."""
    src = os.path.join(_HERE, "gdr_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgdr_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        lib = ctypes.CDLL(_SO)
        fp = ctypes.c_void_p
        lib.gdr_oracle_recurrent_f32.argtypes = [fp] * 8 + [ctypes.c_int] * 5 + [ctypes.c_float, ctypes.c_int]
        lib.gdr_oracle_recurrent_f32.restype = ctypes.c_int
        lib.gdr_oracle_num_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().gdr_oracle_num_threads())


def gdr_recurrent_c(q, k, v, g, beta, scale: Optional[float] = None,
                    initial_state: Optional[torch.Tensor] = None,
                    nthreads: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same contract as oracle.gdr_ref.gdr_recurrent_ref, computed by the plain-C port."""
    lib = _load()
    B, T, H, K = k.shape
    V = v.shape[-1]
    if scale is None:
        scale = 1.0 / math.sqrt(K)
    f = lambda x: x.detach().to("cpu", torch.float32).contiguous()
    q, k, v, g, beta = map(f, (q, k, v, g, beta))
    s0 = f(initial_state) if initial_state is not None else None
    o = torch.empty(B, T, H, V, dtype=torch.float32)
    sT = torch.empty(B, H, K, V, dtype=torch.float32)
    rc = lib.gdr_oracle_recurrent_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(),
                                      beta.data_ptr(), s0.data_ptr() if s0 is not None else None,
                                      o.data_ptr(), sT.data_ptr(), B, T, H, K, V, float(scale), nthreads)
    if rc != 0:
        raise RuntimeError("gdr_oracle_recurrent_f32 failed")
    return o, sT
