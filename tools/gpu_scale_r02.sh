#!/bin/bash
# Multi-GPU pass of round 2: bench.py at N = 1 and N = $1.. (own arm + reference arm), the multi-device handle test.
set -u
mkdir -p gpurun_out
T=${T:-r02h}
python -m pytest tests/test_gpu_device_build.py tests/test_gpu_nontriangle.py -x -q -m gpu 2>&1 | tail -2
for n in "$@"; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/${T}_scale_n$n.json 2> gpurun_out/${T}_scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${T}_scale_n$n.json 2> gpurun_out/${T}_scale_n$n.err
  fi
  echo "N=$n rc=$? lines=$(wc -l < gpurun_out/${T}_scale_n$n.json)"; tail -3 gpurun_out/${T}_scale_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/${T}_scale_n$n.json')); print(d['n_gpus'], 'value', round(d['value']), 'primary', round(d['detail']['primary_mrays_s']), 'bounce', round(d['detail']['bounce_mrays_s']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['detail']['e2e_ms_per_step'],3), 'roofline', round(d['roofline']['frac'],3))"
done
