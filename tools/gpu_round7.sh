#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/check_variants.py 1 0 3 2>&1 | tail -5
MIROGPU_PF=5 MIROGPU_MINB=10 python tools/check_variants.py 1 3 2>&1 | tail -1
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
run "v3 base" MIROGPU_VARIANT=3
for pf in 1 2 4 5 6; do run "v3 pf=$pf" MIROGPU_VARIANT=3 MIROGPU_PF=$pf; done
run "v3 minb=10" MIROGPU_VARIANT=3 MIROGPU_MINB=10
run "v3 minb=12" MIROGPU_VARIANT=3 MIROGPU_MINB=12
run "v3 pf=5 minb=10" MIROGPU_VARIANT=3 MIROGPU_PF=5 MIROGPU_MINB=10
run "v3 pf=5 minb=12" MIROGPU_VARIANT=3 MIROGPU_PF=5 MIROGPU_MINB=12
run "v3 l2persist=1" MIROGPU_VARIANT=3 MIROGPU_L2PERSIST=1
run "v3 l2persist=2" MIROGPU_VARIANT=3 MIROGPU_L2PERSIST=2
run "v3 l2persist=1 pf=5" MIROGPU_VARIANT=3 MIROGPU_L2PERSIST=1 MIROGPU_PF=5
grep -h "L2 persisting" gpurun_out/bench_x.err | head -2
