#!/bin/bash
# hybrid kernel (variant 3): correctness vs variant 0, knob sweep, lane-utilisation metrics
set -u
mkdir -p gpurun_out
python tools/check_variants.py 0 2 3 2>&1 | tail -12
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; return; }
  python - "$label" <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
PY
}
run "v0" MIROGPU_X=0
for nmin in 8 12 16 20; do for period in 2 4 8; do
  run "v3 nmin=$nmin period=$period idle=4" MIROGPU_VARIANT=3 MIROGPU_NMIN=$nmin MIROGPU_PERIOD=$period MIROGPU_MINIDLE=4
done; done
run "v3 nmin=12 period=4 idle=1" MIROGPU_VARIANT=3 MIROGPU_NMIN=12 MIROGPU_PERIOD=4 MIROGPU_MINIDLE=1
run "v3 nmin=12 period=4 idle=8" MIROGPU_VARIANT=3 MIROGPU_NMIN=12 MIROGPU_PERIOD=4 MIROGPU_MINIDLE=8
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
for v in 0 3; do
MIROGPU_VARIANT=$v ncu --metrics $M --clock-control none -k regex:k_trace -s 8 -c 2 --csv --log-file gpurun_out/ncu_v$v.csv python bench.py --no-cpu --steps 2 --warmup 3 > /dev/null 2>&1
python - $v <<'PY'
import csv, sys
rows=[r for r in csv.reader(open(f'gpurun_out/ncu_v{sys.argv[1]}.csv')) if len(r)>10]
hdr=rows[0]; 
for r in rows[1:]:
    d=dict(zip(hdr,r)); print('ncu v'+sys.argv[1], d.get('ID'), d.get('Kernel Name','')[:40], d.get('Metric Name'), d.get('Metric Value'))
PY
done
