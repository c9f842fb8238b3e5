#!/bin/bash
# Round-2 GPU-box pass: tests, smoke, bench (own arm with CPU leg + extras, reference arm), forced variants, then -- only if the
# plain bench exited 0 -- the ncu passes of the same command (launch list incl. the e2e frames, one full capture of the two
# trace launches of a timed step).  Outputs under gpurun_out/ with the tag given as $1 (default r02).
set -u
T=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${T}_gpu_tests.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/${T}_smoke.log)"
python bench.py > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${T}_bench_full.json
python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "bench reference rc=$?"; cut -c1-300 gpurun_out/${T}_bench_reference.json
for v in 0 1; do python bench.py --variant $v --no-cpu --no-extras --steps 10 > gpurun_out/${T}_bench_v$v.json 2> gpurun_out/${T}_bench_v$v.err; echo "bench variant $v rc=$?"; done
# --sequential: whole-batch launches one after the other -- the launches roofline.achieved is measured on (bench.py replays the
# timed steps that way for its per-kernel events); the default timed region issues the same work as half-batches on two streams
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --sequential"
$BENCH > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv $BENCH > gpurun_out/${T}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$BENCH > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 7 -c 2 -f -o gpurun_out/${T}_prof $BENCH > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${T}_prof.ncu-rep --page raw --csv > gpurun_out/${T}_prof.raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_prof.ncu-rep --page source --csv > gpurun_out/${T}_prof.source.csv 2>/dev/null
ls -la gpurun_out | grep ${T}_ | head -40
