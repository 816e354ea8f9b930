#!/bin/bash
# gpurun helper (1 GPU): whole GPU suite, smoke(), the default bench line, then the ncu launch list of a short bench run
# restricted to the library's kernels (every launch of the run: warm-up, timed steps, e2e chunks, MLP epochs).
# COST: the suite, smoke and the two bench lines take ~3 minutes of box time; the launch list took another 16 minutes (5 254
# launches under ncu) -- pass "nolist" as the first argument to skip it.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/full_pytest.log 2>&1
echo "pytest gpu rc=$? $(tail -4 gpurun_out/full_pytest.log | head -1)" | tee gpurun_out/full_summary.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/full_smoke.log 2>&1
echo "smoke rc=$? $(grep 'smoke ok' gpurun_out/full_smoke.log)" | tee -a gpurun_out/full_summary.txt
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/full_bench.err
echo "bench rc=$?" | tee -a gpurun_out/full_summary.txt
( time timeout 600 python bench.py --config c4 ) > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/full_bench_c4.err
echo "bench c4 rc=$?" | tee -a gpurun_out/full_summary.txt
[ "${1:-}" = nolist ] && exit 0
python bench.py --no-cpu --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'extract_kernel|resample_kernel|gemm_|softmax|sgd_|prep_batch|downmix|train_small' -c 8000 --csv --log-file gpurun_out/launches_bench_raw.csv python bench.py --no-cpu --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?" | tee -a gpurun_out/full_summary.txt
