"""Per-phase cycle breakdown of the chunk kernel (CTA 0), from a -DGDKVM_PHASE_TIMERS build.
Build here:  python scripts/phase_timers.py build     Run (GPU box):  python scripts/phase_timers.py [flags]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROF_LIB = os.path.join(ROOT, "gdkvm_b200", "libgdkvm_gdr_prof.so")
NAMES = {0: "K wait(tiles,KQ,buffers)", 1: "K gating", 2: "K gating barrier", 3: "K solve levels 0-1", 4: "K solve level 2",
         5: "K solve barrier", 6: "K T' conv", 7: "K (unused)", 22: "S W^T mma", 23: "S wait O accum", 32: "iS wait K side", 38: "iS issue U", 33: "iS wait W^T", 34: "iS wait Sb + Vn corr", 35: "iS wait Ofree + d1", 36: "iS wait K copy", 37: "iS wait Vnb + e,d2", 24: "S wait staging+bar", 16: "S wait K side", 17: "S wait state upd", 18: "S S-pass",
         19: "S readout", 20: "S wait Vn", 21: "S Vnb pass"}
if len(sys.argv) > 1 and sys.argv[1] == "build":
    from gdkvm_b200 import _build
    cmd = _build.nvcc_command(out=PROF_LIB, extra=["-DGDKVM_PHASE_TIMERS"])
    subprocess.check_call(cmd)
    print("built", PROF_LIB); sys.exit(0)
import torch
from gdkvm_b200 import _build, _cabi
_build.LIB_PATH = _cabi.LIB_PATH = PROF_LIB
import gdkvm_b200
from bench import make_device_inputs
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, T, H, K, V, C = 19, 128 * 49, 8, 64, 256, 49          # 152 chains ~ one wave
q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1, torch.device("cuda"))
lib = _cabi.load()
for it in range(3):
    o, sT = gdkvm_b200.gdr_lkva(q, k, v, g, beta, None, S0, True, C, flags)
    out = (ctypes.c_ulonglong * 64)()
    assert lib.gdkvm_debug_phase_cycles(out, 64) == 0
flat = not (flags & 8)
nchunks = (T + 63) // 64 if flat else (T // C) * ((C + 63) // 64)
tk = sum(out[i] for i in range(0, 8)); ts = sum(out[i] for i in range(16, 25))
print(f"flags={flags} chunks={nchunks}  K-group cycles/chunk {tk / nchunks:.0f}   state-group cycles/chunk {ts / nchunks:.0f}")
for i, name in NAMES.items():
    print(f"  [{i:2d}] {name:28s} {out[i] / nchunks:8.0f} cycles/chunk")

tr = (ctypes.c_longlong * 512)()
if hasattr(lib, "gdkvm_debug_phase_trace") and lib.gdkvm_debug_phase_trace(tr, 512) == 0:
    ev = [(tr[s * 8 + c], s, c) for s in NAMES for c in range(8) if tr[s * 8 + c] > 0]
    if ev:
        t0 = min(e[0] for e in ev)
        print("timeline of CTA 0 (cycles since the first traced event; phase END times; chunk = 40 + c)")
        for t, s, c in sorted(ev):
            role = "K " if s < 8 else ("S " if s < 32 else "iS")
            print(f"  {t - t0:7d}  {role} chunk {40 + c}  end of [{s:2d}] {NAMES[s]}")
