#!/usr/bin/env bash
set -u
TAG=${1:-r2c}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_projection.py tests/test_model_skeleton.py tests/test_gpu_backward.py -m gpu -x -q -s > $OUT/${TAG}_pytest_proj.log 2>&1; echo "proj pytest exit $?"; tail -30 $OUT/${TAG}_pytest_proj.log
timeout 900 python bench.py --no-fla --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
b = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("main", b["ms_per_step"], b["roofline"]["frac"], "sustained", b["sustained"]["ms_per_step"])
for k, v in b["extra"].items():
    print(k, json.dumps(v)[:900])
PY
tail -5 $OUT/${TAG}_bench.err
