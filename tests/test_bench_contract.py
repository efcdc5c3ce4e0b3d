"""CPU tests of bench.py's contract: the reference arm runs here (no GPU) and prints one JSON line with the keys the driver reads;
the helper arithmetic (algorithmic bytes, clock windows) is what DESIGN.md section 5 says."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "gdr_memory_frames_per_s" and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["vs_baseline"] is None
    assert line["config"]["workload"] == "echonet_batch"


def test_algorithmic_bytes_and_clock_windows():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algorithmic_bytes(64, 6272, 8, 64, 256) == 4203216896            # SURVEY.md section 8(d), configs[1]
    assert bench.algorithmic_bytes(1, 1, 1, 64, 256) == 1288 + 2 * 64 * 256 * 4
    s = bench.ClockSampler.__new__(bench.ClockSampler)                             # no nvidia-smi here: feed it rows
    now = time.time() - 10
    rows = [(now + 0.1 * i, 1965.0 if i < 5 else 1650.0, 1965.0, 900.0 + i, {"sw_power_cap"} if i >= 5 else set()) for i in range(10)]
    s._parse = lambda: rows
    w = s.window(now, now + 0.42)
    assert w["samples"] == 5 and w["sm_mhz"] == 1965.0 and w["reasons"] == []
    w = s.window(now + 0.58, now + 0.92)
    assert w["samples"] == 4 and w["sm_mhz"] == 1650.0 and w["reasons"] == ["sw_power_cap"] and w["power_w_max"] == 909.0
    assert s.window(now + 5, now + 6)["samples"] == 0
