#!/bin/bash
# gpurun --gpus N helper: the multi-GPU parity tests, then bench.py at every N given.   usage: tools/gpu_call_multi.sh 2 [4 8]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/multi_pytest.log 2>&1
echo "pytest multi rc=$? $(tail -4 gpurun_out/multi_pytest.log | head -1)" | tee gpurun_out/multi_summary.txt
for n in "$@"; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 ) > gpurun_out/multi_bench_n$n.json 2> gpurun_out/multi_bench_n$n.err
  echo "bench n=$n rc=$?" | tee -a gpurun_out/multi_summary.txt
done
