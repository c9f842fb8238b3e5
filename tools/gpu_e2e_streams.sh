set -u
python -m pytest tests/test_gpu_render.py tests/test_gpu_nontriangle.py -x -q -m gpu 2>&1 | tail -3
for st in 1 2; do
MIROGPU_RENDER_STREAMS=$st python tools/bench_e2e.py 2>&1 | tail -1
done
for sh in 0 8; do
python bench.py --no-cpu --no-extras --steps 40 --warmup 5 --emulate-shard $sh 2> gpurun_out/sched.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('shard $sh', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['detail']['e2e_ms_per_step'],4))"
done
