#!/usr/bin/env bash
set -u
TAG=${1:-r2f}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log
python scripts/bwd_phase_timers.py 2>&1 | tee $OUT/${TAG}_bwd_phases.log
python scripts/time_bwd.py 2>&1 | tee $OUT/${TAG}_time.log
timeout 400 python bench.py --fla-child 2>&1 | tail -1 | tee $OUT/${TAG}_fla.json
