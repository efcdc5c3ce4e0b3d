#!/usr/bin/env bash
# A/B of library builds on configs[1], alternating processes.  Usage (under gpurun): bash scripts/ab.sh "<lib:SEGS> ..." [rounds]
set -u
for r in $(seq 1 ${2:-2}); do
  for spec in $1; do
    lib=${spec%%:*}; segs=${spec##*:}
    if [ "$lib" = default ]; then unset GDKVM_LIB; else export GDKVM_LIB=gdkvm_b200/libgdkvm_gdr_$lib.so; fi
    echo -n "$lib segs=$segs  "
    SEGS=$segs ROUNDS=${ROUNDS:-2} timeout 200 python scripts/seg_sweep.py 64 128 49 8 ${REPS:-30} 2>&1 | tail -1
  done
done
