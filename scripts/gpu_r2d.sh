#!/usr/bin/env bash
set -u
TAG=${1:-r2d}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_projection.py tests/test_model_skeleton.py tests/test_gpu_backward.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/${TAG}_pytest.log
for i in 1 2; do
  python scripts/time_bwd.py
  GDKVM_LIB=gdkvm_b200/libgdkvm_gdr_var_bwd_nopipe.so python scripts/time_bwd.py
done 2>&1 | tee $OUT/${TAG}_time_bwd.log
