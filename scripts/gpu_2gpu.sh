#!/usr/bin/env bash
# Two-GPU visit: the second-device test (one process, two GPUs), then the driver's 2-GPU bench command for both arms.
set -u
TAG=${1:-r2m}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "second_device or sharding" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/${TAG}_bench_2gpu.json 2> $OUT/${TAG}_bench_2gpu.err; echo "bench exit $?"
tail -3 $OUT/${TAG}_bench_2gpu.err
python - <<PY
import json
b = json.loads(open("$OUT/${TAG}_bench_2gpu.json").read().strip().splitlines()[-1])
print("2gpu", b["value"], b["ms_per_step"], "e2e", b["e2e"]["value"], "copy_only", b["e2e"]["copy_only_ms"], "gather", b.get("readout_gather_ms"))
print("sustained", b.get("sustained", {}).get("ms_per_step")); print("extra", json.dumps(b.get("extra"))[:900])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref_2gpu.json 2>> $OUT/${TAG}_bench_2gpu.err; echo "ref exit $?"; cut -c1-200 $OUT/${TAG}_bench_ref_2gpu.json
