#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/check_variants.py 1 2 2>&1 | tail -1
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; return; }
  python - "$label" <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
PY
}
run "v2 default (pf4 minb9 nrep2 pool64)" MIROGPU_VARIANT=2
run "v2 pf=12" MIROGPU_VARIANT=2 MIROGPU_PF=12
run "v2 pf=12 pool=32" MIROGPU_VARIANT=2 MIROGPU_PF=12 MIROGPU_POOL=32
run "v2 pf=0" MIROGPU_VARIANT=2 MIROGPU_PF=0
run "v2 pool=32" MIROGPU_VARIANT=2 MIROGPU_POOL=32
run "v2 pool=128" MIROGPU_VARIANT=2 MIROGPU_POOL=128
run "v2 pool=256" MIROGPU_VARIANT=2 MIROGPU_POOL=256
run "v2 minb8 nrep3" MIROGPU_VARIANT=2 MIROGPU_MINB=8 MIROGPU_NREP=3
run "v2 minb9 nrep3" MIROGPU_VARIANT=2 MIROGPU_MINB=9 MIROGPU_NREP=3
run "v2 nrep2 nmin=12 period=6" MIROGPU_VARIANT=2 MIROGPU_NMIN=12 MIROGPU_PERIOD=6
run "v2 nrep2 nmin=20 period=6" MIROGPU_VARIANT=2 MIROGPU_NMIN=20 MIROGPU_PERIOD=6
run "v2 nrep2 period=4" MIROGPU_VARIANT=2 MIROGPU_PERIOD=4
run "v2 nrep2 period=4 idle=8" MIROGPU_VARIANT=2 MIROGPU_PERIOD=4 MIROGPU_MINIDLE=8
