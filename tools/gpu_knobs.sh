#!/bin/bash
# Knob sweep in one gpurun call: each argument is "label:ENV=VALUE,ENV=VALUE" (empty list = defaults); runs bench.py --no-cpu --no-extras.
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  env $(echo "$envs" | tr ',' ' ') python bench.py --no-cpu --no-extras --steps 20 > gpurun_out/knob_$label.json 2> gpurun_out/knob_$label.err || { echo "$label FAILED"; tail -3 gpurun_out/knob_$label.err; continue; }
  python - "$label" "$envs" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/knob_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
det = d.get("detail", {})
print(sys.argv[1], sys.argv[2], "value", round(d["value"]), "ms_step", round(d["ms_per_step"], 3), "kern", json.dumps({k: round(v, 3) for k, v in det.get("per_gpu_ms", {}).items() if k != "note"}), "e2e", round(d["e2e"]["value"]))
PY
done | tee gpurun_out/knob_summary.txt
