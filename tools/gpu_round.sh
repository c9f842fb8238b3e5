#!/bin/bash
# One GPU-box round: tests, smoke, bench (own arm + reference arm), then -- only if the plain bench exited 0 -- the
# ncu passes of the same command (launch list, one full capture of the trace kernels of a timed step).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/gpu_tests.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench_full.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"; tail -c 1200 gpurun_out/bench_reference.json
for v in 0 1 2; do python bench.py --variant $v --no-cpu --steps 10 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench variant $v rc=$?"; done
python bench.py --layout cwbvh8 --no-cpu --steps 10 > gpurun_out/bench_cwbvh8.json 2> gpurun_out/bench_cwbvh8.err; echo "bench cwbvh8 rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$BENCH > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 7 -c 2 -f -o gpurun_out/prof $BENCH > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof.raw.csv 2>/dev/null
ncu -i gpurun_out/prof.ncu-rep --page source --csv > gpurun_out/prof.source.csv 2>/dev/null
ls -la gpurun_out | head -50
