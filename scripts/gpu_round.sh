#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench (ours + reference arm), then the ncu launch list and
# one full capture of the dominant kernel (each only after the same command exited 0 without ncu).
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag] [bench flags...]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu $*"
timeout 600 $CMD > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
timeout 600 $CMD > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gdr_ -s 3 -c 1 -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu exit $?"; ls -la $OUT
