#!/usr/bin/env bash
# Evidence visit for the shipped build: bench line, ncu launch list of the timed region, one `ncu --set full` capture of each
# hot kernel (chunk kernel from bench.py; backward and projection kernels from scripts/prof_round2.py) -- each only after
# the same command exited 0 without ncu.
set -u
TAG=${1:-r2z}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --no-fla > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras --sustained-seconds 0"
timeout 300 $CMD > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list exit $?"
timeout 300 $CMD > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gdr_chunk_kernel -s 3 -c 1 -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu chunk exit $?"
timeout 300 python scripts/prof_round2.py > $OUT/${TAG}_plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gdr_bwd_kernel -s 1 -c 1 -o $OUT/${TAG}_bwd python scripts/prof_round2.py > $OUT/${TAG}_ncu_bwd.log 2>&1
echo "ncu bwd exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qkvgb_proj_kernel -s 1 -c 1 -o $OUT/${TAG}_proj python scripts/prof_round2.py > $OUT/${TAG}_ncu_proj.log 2>&1
echo "ncu proj exit $?"
ls -la $OUT | grep ${TAG}
