#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/check_variants.py 1 0 2 2>&1 | tail -5
MIROGPU_NREP=2 MIROGPU_MINB=9 MIROGPU_POOL=256 python tools/check_variants.py 1 2 2>&1 | tail -1
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; return; }
  python - "$label" <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
PY
}
run "v0" MIROGPU_VARIANT=0
run "v2 default (pf4 minb10 nrep1 pool64)" MIROGPU_VARIANT=2
run "v2 minb=9" MIROGPU_VARIANT=2 MIROGPU_MINB=9
run "v2 minb=8" MIROGPU_VARIANT=2 MIROGPU_MINB=8
run "v2 nrep=2" MIROGPU_VARIANT=2 MIROGPU_NREP=2
run "v2 nrep=3" MIROGPU_VARIANT=2 MIROGPU_NREP=3
run "v2 minb=9 nrep=2" MIROGPU_VARIANT=2 MIROGPU_MINB=9 MIROGPU_NREP=2
run "v2 minb=8 nrep=2" MIROGPU_VARIANT=2 MIROGPU_MINB=8 MIROGPU_NREP=2
run "v2 pf=0" MIROGPU_VARIANT=2 MIROGPU_PF=0
run "v2 pool=32" MIROGPU_VARIANT=2 MIROGPU_POOL=32
run "v2 pool=128" MIROGPU_VARIANT=2 MIROGPU_POOL=128
run "v2 pool=256" MIROGPU_VARIANT=2 MIROGPU_POOL=256
run "v2 period=4" MIROGPU_VARIANT=2 MIROGPU_PERIOD=4
run "v2 period=12" MIROGPU_VARIANT=2 MIROGPU_PERIOD=12
run "v2 nmin=12" MIROGPU_VARIANT=2 MIROGPU_NMIN=12
run "v2 nmin=20" MIROGPU_VARIANT=2 MIROGPU_NMIN=20
run "v2 idle=8" MIROGPU_VARIANT=2 MIROGPU_MINIDLE=8
run "v2 idle=2" MIROGPU_VARIANT=2 MIROGPU_MINIDLE=2
