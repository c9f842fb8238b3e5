#!/bin/bash
set -u
mkdir -p gpurun_out
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; return; }
  python - "$label" <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
PY
}
for cfg in "16 4 8 0" "16 4 8 4" "16 3 8 0" "16 2 8 0" "16 4 12 0" "16 4 16 0" "12 4 8 0" "14 4 8 0" "18 4 8 0" "16 3 12 0" "12 3 12 0" "14 2 12 0"; do set -- $cfg
  run "v2 nmin=$1 period=$2 idle=$3 pf=$4" MIROGPU_VARIANT=2 MIROGPU_NMIN=$1 MIROGPU_PERIOD=$2 MIROGPU_MINIDLE=$3 MIROGPU_PF=$4
done
run "v2 16 4 8 0 pool32" MIROGPU_VARIANT=2 MIROGPU_NMIN=16 MIROGPU_PERIOD=4 MIROGPU_MINIDLE=8 MIROGPU_PF=0 MIROGPU_POOL=32
run "v2 16 4 8 4 minb8" MIROGPU_VARIANT=2 MIROGPU_NMIN=16 MIROGPU_PERIOD=4 MIROGPU_MINIDLE=8 MIROGPU_PF=4 MIROGPU_MINB=8
