"""Where a work unit of the chunk kernel spends its time outside the steady-state chunk loop (-DGDKVM_PHASE_TIMERS build):
clock64 of CTA 0 at fixed points, one wave of 148 chains, NC chunks per chain.
Build here: python scripts/phase_timers.py build      Run (GPU box): python scripts/unit_marks.py [chunks]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROF_LIB = os.path.join(ROOT, "gdkvm_b200", "libgdkvm_gdr_prof.so")
os.environ["GDKVM_LIB"] = PROF_LIB
import torch
import gdkvm_b200
from gdkvm_b200 import _cabi
from bench import make_device_inputs
dev = torch.device("cuda", 0)
NAMES = ["kernel entry", "setup done (barriers, zeroed tiles, TMEM allocated)", "initial state in TMEM", "K side of the first chunk published",
         "K side of the last chunk published", "", "last Vnb pass done", "last readout drained", "last state update complete",
         "final state stored", "readout stores complete", "teardown barrier passed", "TMEM released"]
lib = _cabi.load()
for nc in [int(a) for a in sys.argv[1:]] or [1, 8, 49]:
    B, T, H, K, V = 37, 64 * nc, 4, 64, 256
    q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234, dev)
    o = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=dev); sT = torch.empty_like(S0)
    for _ in range(3):
        gdkvm_b200.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 0, 0x2 | (1 << 8))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    gdkvm_b200.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, 0, 0x2 | (1 << 8))
    e1.record(); torch.cuda.synchronize()
    m = (ctypes.c_longlong * 16)()
    assert lib.gdkvm_debug_unit_marks(m, 16) == 0
    print(f"{nc} chunks per chain: {e0.elapsed_time(e1) * 1e3:.1f} us between events")
    for i, name in enumerate(NAMES):
        if name and m[i]:
            print(f"   {m[i] - m[0]:8d} cycles  {name}")
