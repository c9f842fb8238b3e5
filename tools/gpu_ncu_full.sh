#!/bin/bash
# one full ncu capture of the bounce-ray trace launch for a kernel variant: tools/gpu_ncu_full.sh <variant> [extra env]
set -u
V=${1:-3}
mkdir -p gpurun_out
MIROGPU_VARIANT=$V ncu --set full --clock-control none --import-source on -k regex:k_trace -s 9 -c 1 -f -o gpurun_out/full_v$V python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_full_v$V.log 2>&1
echo rc=$?
ncu -i gpurun_out/full_v$V.ncu-rep --page raw --csv > gpurun_out/full_v$V.raw.csv 2>/dev/null
ncu -i gpurun_out/full_v$V.ncu-rep --page source --csv > gpurun_out/full_v$V.source.csv 2>/dev/null
ls -la gpurun_out/full_v$V.*
