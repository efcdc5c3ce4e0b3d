#!/usr/bin/env bash
# Build (if needed) and run the tcgen05/TMA layout probes.  Under gpurun: bash scripts/run_probe.sh
set -u
mkdir -p gpurun_out
BIN=tests/probes/umma_probe.bin
if [ ! -x $BIN ] || [ tests/probes/umma_probe.cu -nt $BIN ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O2 -std=c++17 -o $BIN tests/probes/umma_probe.cu || exit 1
fi
if command -v nvidia-smi >/dev/null 2>&1; then timeout 120 $BIN 2>&1 | tee gpurun_out/probe.log; fi
