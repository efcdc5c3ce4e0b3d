"""Time configs[1] with several -D build variants of the library (one .so each).
Build here:  python scripts/variants.py build "NAME:-DX=1 -DY=2" ...     Run (GPU box):  python scripts/variants.py run NAME ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = lambda name: os.path.join(ROOT, "gdkvm_b200", f"libgdkvm_gdr_var_{name}.so")
if sys.argv[1] == "build":
    from gdkvm_b200 import _build
    procs = []
    for spec in sys.argv[2:]:
        name, _, flags = spec.partition(":")
        procs.append((name, subprocess.Popen(_build.nvcc_command(out=lib(name), extra=flags.split()))))
    for name, pr in procs:
        assert pr.wait() == 0, name
    print("built", [n for n, _ in procs]); sys.exit(0)
for name in sys.argv[2:]:
    env = dict(os.environ, GDKVM_LIB=lib(name))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-e2e", "--no-cpu"],
                         capture_output=True, text=True, env=env)
    try:
        ms = json.loads(out.stdout.strip().splitlines()[-1])["ms_per_step"]
        print(f"{name:24s} {ms:7.4f} ms/step", flush=True)
    except Exception:
        print(f"{name:24s} FAILED {out.stderr[-300:]}", flush=True)
