#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/check_variants.py 1 0 2 2>&1 | tail -1
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; return; }
  python - "$label" <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
PY
}
for ppt in 1 2 4 8; do
run "bvh2 v0 ppt=$ppt" MIROGPU_LAYOUT=bvh2 MIROGPU_VARIANT=0 MIROGPU_PPT=$ppt
run "qbvh4 v0 ppt=$ppt" MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=0 MIROGPU_PPT=$ppt
done
run "bvh4 v0 ppt=4" MIROGPU_LAYOUT=bvh4 MIROGPU_VARIANT=0 MIROGPU_PPT=4
run "cwbvh8 ppt=4" MIROGPU_LAYOUT=cwbvh8 MIROGPU_PPT=4
run "bvh2 auto ppt=4" MIROGPU_LAYOUT=bvh2 MIROGPU_PPT=4
run "qbvh4 auto pool128" MIROGPU_LAYOUT=qbvh4 MIROGPU_POOL=128
run "qbvh4 auto pool96" MIROGPU_LAYOUT=qbvh4 MIROGPU_POOL=96
