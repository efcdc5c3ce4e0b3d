#!/usr/bin/env bash
# ncu --set full captures of the backward and projection kernels (one launch each, after a plain run exited 0)
set -u
TAG=${1:-r2n}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/prof_round2.py > $OUT/${TAG}_plain.log 2>&1 || { echo plain run failed; tail -20 $OUT/${TAG}_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gdr_bwd_kernel -s 1 -c 1 -o $OUT/${TAG}_bwd python scripts/prof_round2.py > $OUT/${TAG}_ncu_bwd.log 2>&1; echo "ncu bwd exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qkvgb_proj_kernel -s 1 -c 1 -o $OUT/${TAG}_proj python scripts/prof_round2.py > $OUT/${TAG}_ncu_proj.log 2>&1; echo "ncu proj exit $?"
ls -la $OUT | tail -5
