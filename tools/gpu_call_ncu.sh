#!/bin/bash
# gpurun helper: launch list of one bench.py run (device time per launch) and one `ncu --set full` capture of the kernels of
# two training steps.  Each ncu run follows a plain run of the same command that exited 0.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/gpu_mlp_step.py 3xtf32 6 > gpurun_out/ncu_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tma|softmax_train|sgd_fused|prep_batch' -s 54 -c 18 -o gpurun_out/r02_mlp_tma -f python tools/gpu_mlp_step.py 3xtf32 6 > gpurun_out/ncu_step.log 2>&1
echo "full capture rc=$?"
