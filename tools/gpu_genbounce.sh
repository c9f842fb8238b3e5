#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py tests/test_gpu_photon_trace.py -x -q -m gpu 2>&1 | tail -2
python bench.py --no-cpu --steps 10 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_x.json'))
c = d['config']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'primary', round(c['primary_mrays_s']), 'bounce', round(c['bounce_mrays_s']), 'e2e', round(d['e2e']['value']))
print('trace ms', c['primary_rays_per_step'] / c['primary_mrays_s'] / 1e3, c['bounce_rays_per_step'] / c['bounce_mrays_s'] / 1e3)
PY
ncu --set full --clock-control none --import-source on -k regex:k_gen_bounce -s 3 -c 1 -f -o gpurun_out/genb python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_genb.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/genb.ncu-rep --page raw --csv > gpurun_out/genb.raw.csv 2>/dev/null
ncu -i gpurun_out/genb.ncu-rep --page source --csv > gpurun_out/genb.source.csv 2>/dev/null
rm -f gpurun_out/genb.ncu-rep
