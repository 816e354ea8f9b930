#!/bin/bash
# gpurun helper: the whole GPU test suite, smoke(), then the default bench line (N = 1).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/full_pytest.log 2>&1
echo "pytest gpu rc=$? $(tail -4 gpurun_out/full_pytest.log | head -1)" | tee gpurun_out/full_summary.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/full_smoke.log 2>&1
echo "smoke rc=$? $(grep 'smoke ok' gpurun_out/full_smoke.log)" | tee -a gpurun_out/full_summary.txt
( time timeout 900 python bench.py ) > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
echo "bench rc=$?" | tee -a gpurun_out/full_summary.txt
