#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 2 -f -o gpurun_out/shade python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_shade.log 2>&1
echo rc=$?
ncu -i gpurun_out/shade.ncu-rep --page raw --csv > gpurun_out/shade.raw.csv 2>/dev/null
ncu -i gpurun_out/shade.ncu-rep --page source --csv > gpurun_out/shade.source.csv 2>/dev/null
ls -la gpurun_out/shade.*
