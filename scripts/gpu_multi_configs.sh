#!/bin/bash
# N GPUs of one box: every bench configuration under torchrun (BASELINE configs 2-5 sharded over the ranks), short runs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
port=29540
for args in "--config 1 --steps 20 --warmup 5" "--config 3 --steps 400 --warmup 20" "--config 2 --steps 100 --warmup 10" "--config 4 --steps 40 --warmup 5"; do
  tag=$(echo "$args" | tr -d ' -')
  port=$((port+1))
  ( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --cpu-seconds 0.5 --e2e-steps 3 --no-latency $args ) > gpurun_out/benchcfg_${N}gpu_$tag.log 2>&1
  echo "${N}gpu [$args] rc=$? : $(grep '^{' gpurun_out/benchcfg_${N}gpu_$tag.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('ms/step=%.4f value=%.0f frac=%.3f e2e=%.0f wl=%s par=%s'%(d['ms_per_step'],d['value'],d['roofline']['frac'],d['e2e']['value'],d['config']['workload'][:60],d['config']['parallelism'][:70]))
except Exception as e: print('parse fail', e)
")"
done
