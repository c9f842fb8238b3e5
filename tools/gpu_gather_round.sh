#!/bin/bash
# After a gather change: whole GPU suite, gather throughput (default + exact), per-config frames, one ncu capture of the gather.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/bench_gather.py > gpurun_out/gather.json 2> gpurun_out/gather.err; echo "gather rc=$?"; cat gpurun_out/gather.json; echo
MIRO_GATHER_EXACT=1 python tools/bench_gather.py > gpurun_out/gather_exact.json 2> gpurun_out/gather_exact.err; echo "gather exact rc=$?"
MIRO_REF_ALL=1 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "configs rc=$?"; cat gpurun_out/configs.json; echo
ncu --set full --clock-control none --import-source on -k regex:k_photon_gather_warp -s 2 -c 1 -f -o gpurun_out/gather_full python tools/bench_gather.py > gpurun_out/ncu_gather.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/gather_full.ncu-rep --page raw --csv > gpurun_out/gather_full.raw.csv 2>/dev/null
ncu -i gpurun_out/gather_full.ncu-rep --page source --csv > gpurun_out/gather_full.source.csv 2>/dev/null
rm -f gpurun_out/gather_full.ncu-rep
