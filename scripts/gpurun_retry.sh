#!/usr/bin/env bash
# Local helper (CPU container): call gpurun, retrying while the pod has no free GPU slot (exit code 3 / "transient").
# Usage: scripts/gpurun_retry.sh <log> <timeout-seconds> '<command>' [gpus]
LOG=$1; TO=$2; CMD=$3; GPUS=${4:-1}
for i in $(seq 1 20); do
  if [ "$GPUS" = 1 ]; then /usr/local/graft/bin/gpurun --timeout $TO -- "$CMD" > $LOG 2>&1; else /usr/local/graft/bin/gpurun --gpus $GPUS --timeout $TO -- "$CMD" > $LOG 2>&1; fi
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc = 3 ]; then sleep 150; continue; fi
  break
done
exit $rc
