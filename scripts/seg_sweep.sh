#!/usr/bin/env bash
# Time-segment sweep on configs[1] (and the other bench shapes): forced segment counts vs the library's choice.
# Usage (under gpurun):  bash scripts/seg_sweep.sh <tag>
set -u
TAG=${1:-seg}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "segments or carry or graph" > $OUT/${TAG}_tests.log 2>&1; echo "tests exit $?"; tail -5 $OUT/${TAG}_tests.log
for wl in echonet_batch camus long_clip; do
  for fl in 256 0 512 768 1024; do
    [ "$wl" != echonet_batch ] && [ "$fl" != 256 ] && [ "$fl" != 0 ] && continue
    timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --flags $fl > $OUT/${TAG}_${wl}_$fl.json 2> $OUT/${TAG}_${wl}_$fl.err
    python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_${wl}_$fl.json").read().strip().splitlines()[-1])
    print("$wl flags=$fl", round(b["ms_per_step"], 4), "ms/step  frac", round(b["roofline"]["frac"], 4))
except Exception as e:
    print("$wl flags=$fl unreadable", e); print(open("$OUT/${TAG}_${wl}_$fl.err").read()[-1500:])
PY
  done
done
