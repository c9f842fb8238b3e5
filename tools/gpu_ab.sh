#!/bin/bash
# A/B of prebuilt product libraries in one gpurun call: bash tools/gpu_ab.sh label1:path1 label2:path2 ...   (each run: bench.py --no-cpu --no-extras)
# The library in place at the start is restored at the end.
set -u
mkdir -p gpurun_out
LIB=cse168-raytracer_b200/libmirogpu.so
cp $LIB /tmp/libmirogpu_keep.so
for rep in 1 2; do
for spec in "$@"; do
  label=${spec%%:*}; path=${spec#*:}
  cp "$path" $LIB
  python bench.py --no-cpu --no-extras --steps 20 > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err || { echo "$label FAILED"; tail -3 gpurun_out/ab_$label.err; continue; }
  python - "$label" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/ab_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
det = d.get("detail", {})
print(sys.argv[1], "value", round(d["value"]), "ms_step", round(d["ms_per_step"], 3), "seq_ms", det.get("ms_per_step_sequential"), "kern", json.dumps({k: round(v, 3) for k, v in det.get("per_gpu_ms", {}).items() if k != "note"}), "e2e", round(d["e2e"]["value"]))
PY
done
done | tee gpurun_out/ab_summary.txt
cp /tmp/libmirogpu_keep.so $LIB
