#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; return; }
  python - "$label" <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']), 'nodes', d['config']['nodes'], 'MB', d['config']['node_mb'])
PY
}
run "bvh2 auto" MIROGPU_LAYOUT=bvh2
run "qbvh4 auto (minb9 nrep2)" MIROGPU_LAYOUT=qbvh4
run "qbvh4 v0" MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=0
run "qbvh4 v2" MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=2
run "qbvh4 minb9 nrep1" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1
run "qbvh4 minb8 nrep1" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1 MIROGPU_MINB=8
run "qbvh4 minb8 nrep2" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=2 MIROGPU_MINB=8
run "qbvh4 minb9 nrep1 nmin12" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1 MIROGPU_NMIN=12
run "qbvh4 minb9 nrep1 nmin20" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1 MIROGPU_NMIN=20
run "qbvh4 minb9 nrep1 period2" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1 MIROGPU_PERIOD=2
run "qbvh4 minb9 nrep1 period8" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1 MIROGPU_PERIOD=8
run "qbvh4 minb9 nrep1 idle4" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1 MIROGPU_MINIDLE=4
run "qbvh4 minb9 nrep1 idle12" MIROGPU_LAYOUT=qbvh4 MIROGPU_NREP=1 MIROGPU_MINIDLE=12
