#!/usr/bin/env bash
# A/B timing of projection-kernel library variants (GPU box).  Usage: proj_ab.sh <tag> <variant>...
TAG=$1; shift; OUT=gpurun_out; mkdir -p $OUT
for round in 1 2; do
  for v in default "$@"; do
    L=gdkvm_b200/libgdkvm_gdr_var_$v.so; [ $v = default ] && L=gdkvm_b200/libgdkvm_gdr.so
    echo "$v bias: $(GDKVM_LIB=$L BIAS=1 timeout 120 python scripts/ablate_proj.py time)   no bias: $(GDKVM_LIB=$L BIAS=0 timeout 120 python scripts/ablate_proj.py time)"
  done
done 2>&1 | tee $OUT/${TAG}_proj_ab.log
