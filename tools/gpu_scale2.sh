#!/bin/bash
# N = 2 check of the multi-GPU e2e path (8-bit row gather) + the render tests on one GPU
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py -x -q -m gpu -k "whitted or shards" 2>&1 | tail -2
n=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
echo "N=$n rc=$? lines=$(wc -l < gpurun_out/scale_n$n.json)"; tail -3 gpurun_out/scale_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/scale_n$n.json')); print(d['n_gpus'], 'value', round(d['value']), 'primary', round(d['config']['primary_mrays_s']), 'bounce', round(d['config']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3))"
