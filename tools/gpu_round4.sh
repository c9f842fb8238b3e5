#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "4 1.0" "8 1.0" "8 2.0" "8 3.0" "4 2.0" "2 1.0" "8 0.5" "6 1.5"; do
  set -- $cfg
  MIROGPU_MAX_LEAF=$1 MIROGPU_CTRAV=$2 python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_x.json'))
print('max_leaf=$1 ctrav=$2 value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']), 'nodes', d['config']['nodes'])
PY
done
