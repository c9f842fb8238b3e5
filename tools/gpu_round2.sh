#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/gpu_tests.log)"
for t in 1 2 3 4 6 8 16; do
  MIROGPU_QUANTUM=$t python bench.py --variant 2 --no-cpu --steps 10 > gpurun_out/bench_t$t.json 2>gpurun_out/bench_t$t.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_t$t.json'))
print('quantum=$t value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
PY
done
for v in 0 0; do python bench.py --variant $v --no-cpu --steps 10 > gpurun_out/bench_v$v.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/bench_v$v.json')); print('variant$v value', round(d['value']), round(d['config']['primary_mrays_s']), round(d['config']['bounce_mrays_s']))"; done
