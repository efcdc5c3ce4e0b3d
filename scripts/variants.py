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
    if name == "default":
        env.pop("GDKVM_LIB")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-e2e", "--no-cpu", "--no-extras",
                          "--sustained-seconds", os.environ.get("SUSTAINED", "2.5")], capture_output=True, text=True, env=env)
    try:
        b = json.loads(out.stdout.strip().splitlines()[-1])
        su = b.get("sustained") or {}
        print(f"{name:24s} burst {b['ms_per_step']:7.4f} ms/step   sustained {su.get('ms_per_step', float('nan')):7.4f} "
              f"(last quarter {su.get('last_quarter_ms_per_step', float('nan')):7.4f}) clocks {su.get('clocks')}", flush=True)
    except Exception:
        print(f"{name:24s} FAILED {out.stderr[-300:]}", flush=True)
