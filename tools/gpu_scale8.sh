#!/bin/bash
# bench.py at N = 2, 4, 8 on one 8-GPU box (own arm), and the reference arm under torchrun at N = 8
set -u
mkdir -p gpurun_out
T=${T:-r02k}
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${T}_scale_n$n.json 2> gpurun_out/${T}_scale_n$n.err
  echo "N=$n rc=$? lines=$(wc -l < gpurun_out/${T}_scale_n$n.json)"; tail -2 gpurun_out/${T}_scale_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/${T}_scale_n$n.json')); print(d['n_gpus'], 'value', round(d['value']), 'primary', round(d['detail']['primary_mrays_s']), 'bounce', round(d['detail']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'seq ms', round(d['detail']['ms_per_step_sequential'],3), 'e2e ms', round(d['detail']['e2e_ms_per_step'],3), 'roofline', round(d['roofline']['frac'],3))"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/${T}_ref_n8.json 2> gpurun_out/${T}_ref_n8.err
echo "ref N=8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${T}_ref_n8.json')); print('reference arm under torchrun:', round(d['value'],2), d['detail'].get('threads'), d['cpu_baseline'].get('cores'))"
