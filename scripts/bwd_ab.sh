#!/usr/bin/env bash
# A/B timing of backward-kernel library variants (GPU box): parity of the variant first, then alternating timings.  Usage: bwd_ab.sh <tag> <variant>...
TAG=$1; shift; OUT=gpurun_out; mkdir -p $OUT
for v in "$@"; do
  GDKVM_LIB=gdkvm_b200/libgdkvm_gdr_var_$v.so timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -2
done
for round in 1 2 3; do
  for v in default "$@"; do
    L=gdkvm_b200/libgdkvm_gdr_var_$v.so; [ $v = default ] && L=gdkvm_b200/libgdkvm_gdr.so
    GDKVM_LIB=$L timeout 200 python scripts/time_bwd.py 2>&1 | tail -1 | sed "s|.*libgdkvm_gdr||"
  done
done 2>&1 | tee $OUT/${TAG}_bwd_ab.log
