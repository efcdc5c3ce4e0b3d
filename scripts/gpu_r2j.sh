#!/usr/bin/env bash
set -u
TAG=${1:-r2j}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_projection.py tests/test_model_skeleton.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log
python scripts/time_bwd.py 2>&1 | tee $OUT/${TAG}_time.log
