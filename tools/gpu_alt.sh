#!/bin/bash
# one rank's share of an N-rank run (--emulate-shard) under the default half-batch schedule and under alternating whole steps
set -u
mkdir -p gpurun_out
for sh in 8 2 1; do for alt in 2 3; do
  MIRO_BENCH_INFLIGHT=$alt python bench.py --no-cpu --no-extras --steps 40 --emulate-shard $sh > gpurun_out/alt_${sh}_$alt.json 2> gpurun_out/alt_${sh}_$alt.err || { echo "shard $sh alt $alt FAILED"; tail -3 gpurun_out/alt_${sh}_$alt.err; continue; }
  python - $sh $alt <<'PY'
import json, sys
d = json.loads(open('gpurun_out/alt_%s_%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print("shard", sys.argv[1], "alternate", sys.argv[2], "value", round(d["value"]), "ms_step", round(d["ms_per_step"], 4), "seq_ms", round(d["detail"]["ms_per_step_sequential"], 4))
PY
done; done | tee gpurun_out/alt_summary.txt
