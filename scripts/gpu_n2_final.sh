#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
( timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 scripts/test_peer_gather.py ) > gpurun_out/peer_gather_test_$N.log 2>&1
echo "peer gather test rc=$? : $(grep -E 'PEER_GATHER' gpurun_out/peer_gather_test_$N.log | tail -1)"
port=29570
for args in "--steps 20 --warmup 5" "--impl reference --steps 20 --warmup 5" "--steps 1000 --warmup 20 --no-latency" "--config 5 --steps 20 --warmup 5"; do
  tag=$(echo "$args" | tr -d ' -')
  port=$((port+1))
  ( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --cpu-seconds 1 --e2e-steps 3 $args ) > gpurun_out/benchfin_${N}gpu_$tag.log 2>&1
  echo "${N}gpu [$args] rc=$? : $(grep '^{' gpurun_out/benchfin_${N}gpu_$tag.log | tail -1 | cut -c1-230)"
done
