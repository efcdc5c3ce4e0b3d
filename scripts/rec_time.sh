#!/usr/bin/env bash
# fp32 CUDA-core recurrent kernel at configs[1] (bench.py --flags 1), default build and named variants.  Usage: rec_time.sh <tag> [variant...]
TAG=$1; shift; OUT=gpurun_out; mkdir -p $OUT
for round in 1 2; do
  for v in default "$@"; do
    L=gdkvm_b200/libgdkvm_gdr_var_$v.so; [ $v = default ] && L=gdkvm_b200/libgdkvm_gdr.so
    GDKVM_LIB=$L timeout 300 python bench.py --flags 1 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras --sustained-seconds 0 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'bf16 I/O', round(b['ms_per_step'], 3), 'ms', round(b['value'] / 1e6, 3), 'M frames/s')"
  done
done 2>&1 | tee $OUT/${TAG}_rec_time.log
