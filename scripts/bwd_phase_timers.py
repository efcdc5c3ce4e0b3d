"""Cycles of CTA 0 of the backward kernel between consecutive barriers (-DGDKVM_BWD_TIMERS build).
Build here: python scripts/bwd_phase_timers.py build     Run (GPU box): python scripts/bwd_phase_timers.py"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "gdkvm_b200", "libgdkvm_gdr_var_bwdtimers.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    from gdkvm_b200 import _build
    subprocess.check_call(_build.nvcc_command(out=LIB, extra=["-DGDKVM_BWD_TIMERS"]))
    sys.exit(0)
os.environ["GDKVM_LIB"] = LIB
import torch
import gdkvm_b200
from gdkvm_b200 import _cabi
from bench import make_device_inputs
dev = torch.device("cuda", 0)
B, T, H, K, V = 64, 128 * 49, 8, 64, 256
q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234, dev)
gen = torch.Generator(device=dev).manual_seed(2)
do = torch.randn(B, T, H, V, generator=gen, device=dev).bfloat16()
dsT = torch.randn(B, H, K, V, generator=gen, device=dev)
o, sT, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
lib = _cabi.load()
buf = (ctypes.c_ulonglong * 64)()
torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, True)
lib.gdkvm_debug_bwd_cycles(buf, 64)          # reset after the warm-up
torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, do, dsT, 0.125, True)
lib.gdkvm_debug_bwd_cycles(buf, 64)
names = {0: "chunk top: loads landed, gates", 1: "K K^T, Q K^T -> A, KKD, P; Qe", 2: "solve levels 0-1", 3: "solve level 2", 4: "T -> bf16 tiles",
         5: "W", 6: "(half 1: loads landed)", 7: "dS' copy, <dS', S>", 8: "Vn, dVn", 9: "dS, dP', G2, dQS, dW, dKh, dV",
         10: "dv store / loads issued; dW tile, dP' epilogue", 11: "dQ, dBt, dT", 12: "X = T^T dT", 13: "dA, M", 14: "dK", 15: "dq, dk, dg, dbeta stores"}
nchunks = (T + 63) // 64
tot = sum(buf)
print(f"backward kernel, CTA 0, {nchunks} chunks: {tot / nchunks:.0f} cycles per chunk")
for i in range(64):
    if buf[i]:
        base = i % 24
        print(f"  slot {i:2d} {'half 1 ' if i >= 24 else '       '}{names.get(base, '?'):52s} {buf[i] / nchunks:8.0f} cycles/chunk  {100 * buf[i] / tot:5.1f} %")
