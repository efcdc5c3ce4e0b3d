"""In-tree build of the sm_100a C-ABI library (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
# GDKVM_LIB: load another build of the same ABI (profiling / ablation builds made by scripts/)
LIB_PATH = os.environ.get("GDKVM_LIB") or os.path.join(PKG_DIR, "libgdkvm_gdr.so")
SOURCES = ["gdr_api.cu", "gdr_recurrent.cu", "gdr_chunked_sm100.cu", "gdr_bwd_sm100.cu", "gdr_proj_sm100.cu", "l2norm.cu"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def nvcc_command(out: str = LIB_PATH, extra: List[str] | None = None) -> List[str]:
    return [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
        *(extra or []), "-o", out, *[os.path.join(CSRC, s) for s in SOURCES],
    ]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "gdkvm_gdr.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into gdkvm_b200/libgdkvm_gdr.so."""
    if os.environ.get("GDKVM_LIB"):
        return LIB_PATH
    if force or needs_build():
        cmd = nvcc_command(extra=["-Xptxas", "-v"] if verbose else None)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stderr[-4000:])
    return LIB_PATH
