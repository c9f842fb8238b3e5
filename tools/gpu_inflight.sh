#!/bin/bash
# One rank's share of an N-rank run on one GPU (bench.py --emulate-shard N) per schedule of the timed region:
# MIRO_BENCH_INFLIGHT = 0 (two half-batches per step on two streams), 2, 3 (whole steps in flight).  -> profiles/r02v_inflight.txt
set -u
mkdir -p gpurun_out
for sh in ${SHARDS:-8 4 2 1}; do for k in ${INFLIGHT:-0 2 3}; do
  MIRO_BENCH_INFLIGHT=$k python bench.py --no-cpu --no-extras --steps 40 --emulate-shard $sh > gpurun_out/inflight_${sh}_$k.json 2> gpurun_out/inflight_${sh}_$k.err || { echo "shard $sh in flight $k FAILED"; tail -3 gpurun_out/inflight_${sh}_$k.err; continue; }
  python - $sh $k <<'PY'
import json, sys
d = json.loads(open('gpurun_out/inflight_%s_%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print("shard", sys.argv[1], "steps in flight", sys.argv[2], "value", round(d["value"]), "ms_step", round(d["ms_per_step"], 4), "seq_ms", round(d["detail"]["ms_per_step_sequential"], 4))
PY
done; done | tee gpurun_out/inflight_summary.txt
