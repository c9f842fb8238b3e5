#!/bin/bash
# bench.py at N = $1 (own arm) with the round's final schedule; prints value / e2e.  usage: gpurun --gpus N -- bash tools/gpu_scale_v.sh N [steps]
set -u
mkdir -p gpurun_out
n=$1; K=${2:-20}; T=${T:-r02v}
if [ "$n" = "1" ]; then
  python bench.py --gpus 1 --steps $K --warmup 5 --no-cpu --no-extras > gpurun_out/${T}_scale_n$n.json 2> gpurun_out/${T}_scale_n$n.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps $K --warmup 5 > gpurun_out/${T}_scale_n$n.json 2> gpurun_out/${T}_scale_n$n.err
fi
echo "N=$n rc=$?"; tail -2 gpurun_out/${T}_scale_n$n.err
python -c "
import json; d=json.loads(open('gpurun_out/${T}_scale_n$n.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['detail']['e2e_ms_per_step'],3), 'roofline', round(d['roofline']['frac'],3))"
