#!/usr/bin/env bash
# Correctness read-out + A/B timing of one library variant against the default build.  Usage: gpu_var.sh <variant> [tag]
set -u
VAR=$1; TAG=${2:-var}; OUT=gpurun_out; mkdir -p $OUT
GDKVM_LIB=gdkvm_b200/libgdkvm_gdr_var_$VAR.so timeout 300 python scripts/dbg_chunk.py > $OUT/${TAG}_dbg.log 2>&1; echo "dbg exit $?"; tail -6 $OUT/${TAG}_dbg.log
GDKVM_LIB=gdkvm_b200/libgdkvm_gdr_var_$VAR.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
SUSTAINED=1.0 python scripts/variants.py run default $VAR default $VAR default $VAR 2>&1 | tee $OUT/${TAG}_ab.log | cut -c1-110
