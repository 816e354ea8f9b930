#!/bin/bash
# gpurun helper (1 GPU): whole GPU suite, smoke(), the default bench line and the reference arm, and the other BASELINE configs
# (c1, c4, c5).  ~5 minutes of box time.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/full_pytest.log 2>&1
echo "pytest gpu rc=$? $(tail -4 gpurun_out/full_pytest.log | head -1)" | tee gpurun_out/full_summary.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/full_smoke.log 2>&1
echo "smoke rc=$? $(grep 'smoke ok' gpurun_out/full_smoke.log)" | tee -a gpurun_out/full_summary.txt
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/full_bench.err
echo "bench rc=$?" | tee -a gpurun_out/full_summary.txt
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/full_bench_ref.err
echo "bench reference rc=$?" | tee -a gpurun_out/full_summary.txt
for c in c1 c4 c5; do
  ( time timeout 900 python bench.py --config $c ) > gpurun_out/r02_bench_${c}_n1.json 2> gpurun_out/full_bench_$c.err
  echo "bench $c rc=$?" | tee -a gpurun_out/full_summary.txt
done
