"""Cycles of CTA 0's first epilogue warp and MMA issuer of the projection kernel between consecutive points (-DGDKVM_PROJ_TIMERS).
Build here: python scripts/proj_phase_timers.py build     Run (GPU box): python scripts/proj_phase_timers.py"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "gdkvm_b200", "libgdkvm_gdr_var_projtimers.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    from gdkvm_b200 import _build
    subprocess.check_call(_build.nvcc_command(out=LIB, extra=["-DGDKVM_PROJ_TIMERS"] + sys.argv[2:]))
    sys.exit(0)
os.environ["GDKVM_LIB"] = LIB
import torch
import gdkvm_b200
from gdkvm_b200 import _cabi
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(2)
R = 64 * 6272
x = torch.randn(R, 256, generator=gen, device=dev).bfloat16()
w = (torch.randn(8 * 384 + 16, 256, generator=gen, device=dev) / 16).bfloat16()
b = torch.randn(8 * 384 + 16, generator=gen, device=dev) if os.environ.get("BIAS", "1") == "1" else None
lib = _cabi.load()
buf = (ctypes.c_ulonglong * 32)()
gdkvm_b200.qkvgb_project(x, w, b, 8, 64, 256)
lib.gdkvm_debug_proj_cycles(buf, 32)          # reset after the warm-up
gdkvm_b200.qkvgb_project(x, w, b, 8, 64, 256)
lib.gdkvm_debug_proj_cycles(buf, 32)
bm = 128 if os.environ.get("GDKVM_PROJ_TILE_ROWS") == "128" else 256     # the library picks 256-row tiles at this size
tiles = -(-(-(-R // bm)) // 148) * (13 if bm == 128 else 25)            # output tiles CTA 0 went through
names = {0: "wait: accumulators of the tile complete", 1: "TMEM load (64 columns)", 2: "bias, sum of squares, rsqrt", 3: "wait: staging tile read by the last store",
         4: "scale, pack, staging writes", 5: "proxy fence + warp sync", 6: "bulk store issued", 16: "issuer wait: feature block", 17: "issuer wait: accumulator free",
         18: "issuer wait: weight k-block landed", 19: "issuer: four MMAs + commit"}
for lo, hi, who in ((0, 16, "epilogue warp 2"), (16, 32, "MMA issuer")):
    tot = sum(buf[lo:hi])
    print(f"{who}: {tot / tiles:.0f} cycles per tile ({tiles} tiles)")
    for i in range(lo, hi):
        if buf[i]:
            print(f"  slot {i:2d} {names.get(i, '?'):48s} {buf[i] / tiles:8.0f} cycles/tile  {100 * buf[i] / tot:5.1f} %")
