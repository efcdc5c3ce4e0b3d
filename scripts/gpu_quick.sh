#!/usr/bin/env bash
# Short GPU-box visit while iterating on the chunk kernel: parity read-out, per-phase cycles, device-timed bench.
# Usage (under gpurun):  bash scripts/gpu_quick.sh <tag> [bench flags...]
set -u
TAG=${1:-q}; shift || true
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/dbg_chunk.py > $OUT/${TAG}_dbg.log 2>&1; echo "dbg exit $?"; tail -22 $OUT/${TAG}_dbg.log
timeout 200 python scripts/phase_timers.py > $OUT/${TAG}_phase.log 2>&1; cat $OUT/${TAG}_phase.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("BENCH", b["ms_per_step"], "ms/step", b["value"], "frames/s  frac", b["roofline"]["frac"])
except Exception as e:
    print("bench line unreadable", e); print(open("$OUT/${TAG}_bench.err").read()[-2000:])
PY
