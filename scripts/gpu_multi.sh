#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
( timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/test_peer_gather.py ) > gpurun_out/peer_gather_test_$N.log 2>&1
echo "peer gather test rc=$? : $(grep -E 'PEER_GATHER' gpurun_out/peer_gather_test_$N.log | tail -1)"
port=29530
for args in "" "--collective nccl"; do
  tag=$(echo "$args" | tr -d ' ')
  port=$((port+1))
  ( timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 1000 --warmup 20 --cpu-seconds 1 --e2e-steps 3 --no-latency $args ) > gpurun_out/bench_${N}gpu_$tag.log 2>&1
  echo "${N}gpu [$args] rc=$? : $(grep '^{' gpurun_out/bench_${N}gpu_$tag.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f e2e=%.0f par=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['e2e']['value'],d['config']['parallelism'][:70]))
except Exception as e: print('parse fail', e)
")"
done
