#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
for args in "--streams 1" "--streams 2"; do
  tag=$(echo "$args" | tr -d ' ')
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1001 --warmup 10 --cpu-seconds 1 --e2e-steps 3 $args ) > gpurun_out/bench_2gpu_$tag.log 2>&1
  echo "2gpu [$args] rc=$? : $(grep '^{' gpurun_out/bench_2gpu_$tag.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f e2e=%.0f launch=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['e2e']['value'],d['config']['launch']))
except Exception as e: print('parse fail', e)
")"
  tail -3 gpurun_out/bench_2gpu_$tag.log | cut -c1-300
done
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 ) 2>&1 | tail -1 | cut -c1-300
( timeout 300 python bench.py --steps 1001 --warmup 10 --cpu-seconds 1 --e2e-steps 3 --streams 2 ) 2>&1 | tail -1 | cut -c1-200
