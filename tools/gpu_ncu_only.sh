#!/bin/bash
# the two ncu passes of tools/gpu_round2.sh alone (tag $1), plus the photon tests touched since
set -u
T=${1:-r02n}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_photon_trace.py tests/test_gpu_photon_build.py -x -q -m gpu 2>&1 | tail -2
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --sequential"
$BENCH > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv $BENCH > gpurun_out/${T}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$BENCH > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 7 -c 2 -f -o gpurun_out/${T}_prof $BENCH > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${T}_prof.ncu-rep --page raw --csv > gpurun_out/${T}_prof.raw.csv 2>/dev/null
rm -f gpurun_out/${T}_prof.ncu-rep
