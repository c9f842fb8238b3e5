#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; return; }
  python - "$label" <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
PY
}
run "qbvh4 default" MIROGPU_LAYOUT=qbvh4
run "qbvh4 default again" MIROGPU_LAYOUT=qbvh4
run "qbvh4 pool128" MIROGPU_LAYOUT=qbvh4 MIROGPU_POOL=128
run "qbvh4 pool32" MIROGPU_LAYOUT=qbvh4 MIROGPU_POOL=32
run "qbvh4 nrep3" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=3
run "qbvh4 nrep1" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1
run "qbvh4 minb8" MIROGPU_LAYOUT=qbvh4 MIROGPU_MINB=8
run "bvh2 default" MIROGPU_LAYOUT=bvh2
run "bvh4 default" MIROGPU_LAYOUT=bvh4
