#!/usr/bin/env bash
# GPU visit for the backward pass: its tests (verbose errors), then optional variants A/B.
set -u
TAG=${1:-bwd}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -s > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -60 $OUT/${TAG}_pytest.log
if [ -n "${VARIANTS:-}" ]; then
  timeout 600 python scripts/variants.py run $VARIANTS > $OUT/${TAG}_variants.log 2>&1; cat $OUT/${TAG}_variants.log
fi
