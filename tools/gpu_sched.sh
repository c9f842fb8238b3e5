set -u
for sh in 0 8; do for cfg in "2 2" "4 2" "4 4" "8 2" "8 4" "3 3"; do
set -- $cfg
[ "$1" = "3" ] && continue
MIRO_BENCH_CHUNKS=$1 MIRO_BENCH_STREAMS=$2 python bench.py --no-cpu --no-extras --steps 40 --warmup 5 --emulate-shard $sh 2> gpurun_out/sched.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('shard $sh chunks $1 streams $2', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'seq ms', round(d['detail']['ms_per_step_sequential'],4))"
done; done
tail -3 gpurun_out/sched.err
