#!/bin/bash
# One gpurun call: MLP parity (new fused step structure), step time per SZB_STEP_FUSE setting, then the default bench line.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/f_gpu.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_tmem_a.py -x -q ) > gpurun_out/f_pytest_mlp.log 2>&1
echo "pytest mlp rc=$?" | tee -a gpurun_out/f_summary.txt
for f in 0 1 2 4 6 7; do
  SZB_STEP_FUSE=$f timeout 120 python tools/gpu_mlp_step.py 3xtf32 60 2>&1 | tail -1 | sed "s/^/fuse=$f /" | tee -a gpurun_out/f_summary.txt
done
SZB_STEP_FUSE=7 timeout 120 python tools/gpu_mlp_step.py tf32 60 2>&1 | tail -1 | sed "s/^/fuse=7 /" | tee -a gpurun_out/f_summary.txt
SZB_STEP_FUSE=0 timeout 120 python tools/gpu_mlp_step.py tf32 60 2>&1 | tail -1 | sed "s/^/fuse=0 /" | tee -a gpurun_out/f_summary.txt
