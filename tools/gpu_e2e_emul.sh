#!/bin/bash
# e2e path of one rank's share of an N = 8 run on one GPU (frames in flight, 1-rank NCCL group): library halves on / off, handles 1 / 3
set -u
mkdir -p gpurun_out
for spec in "h3s2:MIRO_BENCH_E2E_HANDLES=3,MIROGPU_RENDER_STREAMS=2" "h3s1:MIRO_BENCH_E2E_HANDLES=3,MIROGPU_RENDER_STREAMS=1" "h4s1:MIRO_BENCH_E2E_HANDLES=4,MIROGPU_RENDER_STREAMS=1" "h2s1:MIRO_BENCH_E2E_HANDLES=2,MIROGPU_RENDER_STREAMS=1"; do
  label=${spec%%:*}; envs=${spec#*:}
  env MIRO_BENCH_FORCE_PIPE=1 $(echo "$envs" | tr ',' ' ') python bench.py --no-cpu --no-extras --steps 40 --emulate-shard 8 > gpurun_out/e2em_$label.json 2> gpurun_out/e2em_$label.err || { echo "$label FAILED"; tail -3 gpurun_out/e2em_$label.err; continue; }
  python - "$label" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/e2em_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["detail"]["e2e_ms_per_step"], 4))
PY
done | tee gpurun_out/e2em_summary.txt
