#!/usr/bin/env bash
# Evidence visit for the projection kernel: plain run, then one `ncu --set full` capture at configs[1] geometry (scripts/prof_proj.py).
set -u
TAG=${1:-r4p}; OUT=gpurun_out; mkdir -p $OUT
timeout 200 python scripts/prof_proj.py > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qkvgb_proj_kernel -s 1 -c 1 -o $OUT/${TAG}_proj python scripts/prof_proj.py > $OUT/${TAG}_ncu_proj.log 2>&1
echo "ncu proj exit $?"
timeout 100 python scripts/write_bw.py | tee $OUT/${TAG}_write_bw.log
