#!/usr/bin/env bash
# Round-2 GPU-box visit: parity tests, smoke, bench with extras (ours + reference arm).  ncu is a separate script.
# Usage (under gpurun):  bash scripts/gpu_round2.sh <tag> [pytest -k expr]
set -u
TAG=${1:-r2}; KEXPR=${2:-}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -s -k "$KEXPR" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
else
  timeout 1800 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
fi
tail -25 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
tail -8 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cut -c1-400 $OUT/${TAG}_bench_ref.json
