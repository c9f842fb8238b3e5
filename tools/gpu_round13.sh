#!/bin/bash
set -u
mkdir -p gpurun_out
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err || { echo "$label FAILED"; tail -3 gpurun_out/bench_x.err; return; }
  python - "$label" <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_x.json'))
print(sys.argv[1], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
PY
}
for cfg in "16 4 8 2" "16 3 8 2" "16 2 8 2" "14 3 8 2" "18 3 8 2" "16 3 6 2" "16 3 10 2" "16 2 6 2" "16 3 8 3" "12 2 8 2"; do set -- $cfg
  run "qbvh4 v2 nmin=$1 period=$2 idle=$3 nrep=$4" MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=2 MIROGPU_NMIN=$1 MIROGPU_PERIOD=$2 MIROGPU_MINIDLE=$3 MIROGPU_NREP=$4
done
run "qbvh4 v2 pf4" MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=2 MIROGPU_PF=4
run "qbvh4 v2 maxleaf8" MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=2 MIROGPU_MAX_LEAF=8
run "qbvh4 v2 maxleaf2" MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=2 MIROGPU_MAX_LEAF=2
run "qbvh4 v2 pool32" MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=2 MIROGPU_POOL=32
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum
MIROGPU_LAYOUT=qbvh4 MIROGPU_VARIANT=2 ncu --metrics $M --clock-control none -k regex:k_trace -s 8 -c 2 --csv --log-file gpurun_out/ncu_q4.csv python bench.py --no-cpu --steps 2 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_q4.csv')) if len(r)>10]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r)); print('ncu q4', d.get('ID'), d.get('Metric Name'), d.get('Metric Value'))
PY
