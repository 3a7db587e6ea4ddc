#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
( timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/test_peer_gather.py ) > gpurun_out/peer_gather_test_$N.log 2>&1
echo "peer gather test rc=$? : $(grep -E 'PEER_GATHER' gpurun_out/peer_gather_test_$N.log | tail -1)"
( timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 1001 --warmup 10 --cpu-seconds 1 --e2e-steps 3 ) > gpurun_out/bench_${N}gpu.log 2>&1
echo "${N}gpu rc=$? : $(grep '^{' gpurun_out/bench_${N}gpu.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f e2e=%.0f par=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['e2e']['value'],d['config']['parallelism'][:70]))
except Exception as e: print('parse fail', e)
")"
grep '^{' gpurun_out/bench_${N}gpu.log | tail -1 > gpurun_out/bench_${N}gpu.json
