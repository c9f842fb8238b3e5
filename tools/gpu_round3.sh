#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/gpu_tests.log)"; grep -E "^E " gpurun_out/gpu_tests.log | head -20
python bench.py --no-cpu --steps 10 > gpurun_out/bench_a.json 2>gpurun_out/bench_a.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_a.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_a.json'))
print('value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']), 'ms/step', d['ms_per_step'])
PY
