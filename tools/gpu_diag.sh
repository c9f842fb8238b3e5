#!/bin/bash
# Short GPU call: the render tests, the small-frame time breakdown, and the per-config frame table.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py -x -q -m gpu > gpurun_out/pytest_render.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_render.log
python tools/diag_frame.py > gpurun_out/diag_frame.json 2> gpurun_out/diag_frame.err; echo "diag rc=$?"; cat gpurun_out/diag_frame.json; tail -3 gpurun_out/diag_frame.err
MIRO_REF_ALL=1 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "configs rc=$?"; cat gpurun_out/configs.json; tail -3 gpurun_out/configs.err
