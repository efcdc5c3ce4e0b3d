"""Ablation timing of the chunk kernel: one library per removed phase (results are WRONG by construction;
only ms/step is read).  Build here:  python scripts/ablate.py build     Run (GPU box):  python scripts/ablate.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NAMES = ["gating", "solve L0-1", "solve L2", "T' conversion", "W^T", "S pass", "Vnb pass", "readout"]
MASKS = [0] + [int(x, 16) for x in os.environ.get("MASKS", "01 02 04 08 10 20 40 80 0f 1f f0 e0 ff").split()]
lib = lambda m: os.path.join(ROOT, "gdkvm_b200", f"libgdkvm_gdr_abl{m:02x}.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    from gdkvm_b200 import _build
    procs = [(m, subprocess.Popen(_build.nvcc_command(out=lib(m), extra=[f"-DGDKVM_ABLATE={m}"]))) for m in MASKS if m]
    for m, pr in procs:
        assert pr.wait() == 0, m
    print("built", len(procs)); sys.exit(0)
for m in MASKS:
    env = dict(os.environ)
    if m:
        env["GDKVM_LIB"] = lib(m)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-e2e", "--no-cpu", "--no-extras", "--sustained-seconds", "0"],
                         capture_output=True, text=True, env=env)
    try:
        ms = json.loads(out.stdout.strip().splitlines()[-1])["ms_per_step"]
    except Exception:
        print(f"mask {m:02x}: FAILED", out.stderr[-300:]); continue
    what = "baseline" if m == 0 else " + ".join(NAMES[i] for i in range(8) if m >> i & 1)
    print(f"mask {m:02x}: {ms:7.4f} ms/step   without: {what}", flush=True)
