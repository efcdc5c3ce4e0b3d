#!/usr/bin/env bash
# N-GPU visit: the driver's bench command for both arms.  Usage: gpu_ngpu.sh <N> <tag>
set -u
N=$1; TAG=${2:-rng}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err; echo "bench exit $?"
tail -3 $OUT/${TAG}_bench_${N}gpu.err
python - <<PY
import json
b = json.loads(open("$OUT/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N", b["n_gpus"], "value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], "copy_only", b["e2e"]["copy_only_ms"], "e2e ms", b["e2e"]["ms_per_step"], "gather", b.get("readout_gather_ms"))
print("sustained", b.get("sustained", {}).get("ms_per_step")); print("extra", json.dumps(b.get("extra"))[:700])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref_${N}gpu.json 2>> $OUT/${TAG}_bench_${N}gpu.err; echo "ref exit $?"; cut -c1-160 $OUT/${TAG}_bench_ref_${N}gpu.json
