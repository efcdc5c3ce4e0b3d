"""Ablation builds of the projection kernel (GDKVM_PROJ_ABLATE bit mask: a phase's work removed, the barrier skeleton intact).
Build here:  python scripts/ablate_proj.py build      Run (GPU box):  python scripts/ablate_proj.py run"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MASKS = [int(m, 0) for m in os.environ.get("MASKS", "0 1 2 4 3 5 8 9 11").split()]
NAMES = {1: "epilogue work", 2: "weight reloads", 4: "global stores", 8: "MMAs"}
lib = lambda m: os.path.join(ROOT, "gdkvm_b200", f"libgdkvm_gdr_var_pabl{m}.so")
if sys.argv[1] == "build":
    from gdkvm_b200 import _build
    procs = [(m, subprocess.Popen(_build.nvcc_command(out=lib(m), extra=[f"-DGDKVM_PROJ_ABLATE={m}"]))) for m in MASKS]
    for m, pr in procs:
        assert pr.wait() == 0, m
    print("built", MASKS); sys.exit(0)
if sys.argv[1] == "run":
    for m in MASKS:
        out = subprocess.run([sys.executable, __file__, "time"], capture_output=True, text=True, env=dict(os.environ, GDKVM_LIB=lib(m)))
        what = " + ".join(v for k, v in NAMES.items() if m & k) or "baseline"
        print(f"mask {m:2d}: {out.stdout.strip() or out.stderr[-300:]}   without: {what}", flush=True)
    sys.exit(0)
import torch
import gdkvm_b200
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(2)
x = torch.randn(64 * 6272, 256, generator=gen, device=dev).bfloat16()
w = (torch.randn(8 * 384 + 16, 256, generator=gen, device=dev) / 16).bfloat16()
b = torch.randn(8 * 384 + 16, generator=gen, device=dev) if os.environ.get("BIAS", "1") == "1" else None
for _ in range(5):
    gdkvm_b200.qkvgb_project(x, w, b, 8, 64, 256)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        gdkvm_b200.qkvgb_project(x, w, b, 8, 64, 256)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 10)
print(f"{min(ts):.4f} ms (median {sorted(ts)[2]:.4f})")
