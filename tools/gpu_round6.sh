#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/check_variants.py 1 0 2 3 2>&1 | tail -13
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
run "v1" MIROGPU_VARIANT=1
run "v2" MIROGPU_VARIANT=2
for cfg in "12 8 4" "16 8 4" "16 8 8" "16 16 8" "12 16 8" "20 8 8"; do set -- $cfg
  run "v3 nmin=$1 period=$2 idle=$3" MIROGPU_VARIANT=3 MIROGPU_NMIN=$1 MIROGPU_PERIOD=$2 MIROGPU_MINIDLE=$3
done
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
for v in 0 3; do
MIROGPU_VARIANT=$v MIROGPU_NMIN=16 MIROGPU_PERIOD=8 MIROGPU_MINIDLE=8 ncu --metrics $M --clock-control none -k regex:k_trace -s 8 -c 2 --csv --log-file gpurun_out/ncu_v$v.csv python bench.py --no-cpu --steps 2 --warmup 3 > /dev/null 2>&1
python - $v <<'PY'
import csv, sys
rows=[r for r in csv.reader(open(f'gpurun_out/ncu_v{sys.argv[1]}.csv')) if len(r)>10]
hdr=rows[0]; 
for r in rows[1:]:
    d=dict(zip(hdr,r)); print('ncu v'+sys.argv[1], d.get('ID'), d.get('Metric Name'), d.get('Metric Value'))
PY
done
