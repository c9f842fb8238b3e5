#!/bin/bash
# Knob sweep helper for one gpurun call: each line of stdin is "label ENV=VALUE ..." and runs bench.py --no-cpu once.
#   echo "qbvh4-nmin12 MIROGPU_LAYOUT=qbvh4 MIROGPU_NMIN=12" | bash tools/gpu_sweep.sh
# Knobs (read by mirogpu_scene_create / bench.py): MIROGPU_LAYOUT bvh2|bvh4|qbvh4|cwbvh8, MIROGPU_VARIANT -1|0|1|2, MIROGPU_BUILDER sah|lbvh,
# MIROGPU_NMIN, MIROGPU_PERIOD, MIROGPU_MINIDLE, MIROGPU_NREP, MIROGPU_POOL, MIROGPU_MINB, MIROGPU_PF, MIROGPU_PPT, MIROGPU_MAX_LEAF, MIROGPU_CTRAV.
set -u
mkdir -p gpurun_out
while read -r label rest; do
  [ -z "$label" ] && continue
  env $rest python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; continue; }
  python - "$label" <<'PY'
import json, sys
d = json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], "value", round(d["value"]), "ms_step", round(d["ms_per_step"], 3), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']),
      'e2e', round(d['e2e']['value']), 'nodes', d['config']['nodes'], 'build_s', round(d['config']['build_s'], 3))
PY
done
