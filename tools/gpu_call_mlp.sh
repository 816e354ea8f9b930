#!/bin/bash
# gpurun helper: MLP parity tests, then the batch-4096 step time for a list of "ENV=VAL,ENV=VAL" settings.
# usage: tools/gpu_call_mlp.sh [--notest] "SZB_STEP_FUSE=4,SZB_GEMM_TMA=0" "SZB_STEP_FUSE=4" ...
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "${1:-}" = "--notest" ]; then shift; else
  ( time timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_tmem_a.py -x -q ) > gpurun_out/pytest_mlp.log 2>&1
  echo "pytest mlp rc=$? $(tail -4 gpurun_out/pytest_mlp.log | head -1)" | tee -a gpurun_out/mlp_summary.txt
fi
for cfg in "$@"; do
  env $(echo "$cfg" | tr ',' ' ') timeout 120 python tools/gpu_mlp_step.py 3xtf32 60 2>&1 | tail -1 | sed "s/^/[$cfg] /" | tee -a gpurun_out/mlp_summary.txt
done
