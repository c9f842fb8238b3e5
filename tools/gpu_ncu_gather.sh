#!/bin/bash
# Photon gather: throughput run, then one full ncu capture of a warp-per-query launch on the global map (config 5).
set -u
mkdir -p gpurun_out
python tools/bench_gather.py > gpurun_out/gather.json 2> gpurun_out/gather.err; echo "gather rc=$?"; cat gpurun_out/gather.json
ncu --set full --clock-control none --import-source on -k regex:k_photon_gather_warp -s 2 -c 1 -f -o gpurun_out/gather_full python tools/bench_gather.py > gpurun_out/ncu_gather.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/gather_full.ncu-rep --page raw --csv > gpurun_out/gather_full.raw.csv 2>/dev/null
ncu -i gpurun_out/gather_full.ncu-rep --page source --csv > gpurun_out/gather_full.source.csv 2>/dev/null
ls -la gpurun_out/gather_full.*
