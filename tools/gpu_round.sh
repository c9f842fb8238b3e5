#!/bin/bash
# One GPU-box round: tests, bench, then (only if the plain bench exited 0) the ncu passes of the same command.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/gpu_tests.log)"
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_full.json
python bench.py --layout cwbvh8 --no-cpu --steps 10 > gpurun_out/bench_cwbvh8.json 2> gpurun_out/bench_cwbvh8.err; echo "bench cwbvh8 rc=$?"
python bench.py --variant 1 --no-cpu --steps 10 > gpurun_out/bench_variant1.json 2> gpurun_out/bench_variant1.err; echo "bench v1 rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu"
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$BENCH > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 7 -c 2 -o gpurun_out/prof $BENCH > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
