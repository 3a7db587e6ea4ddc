#!/bin/bash
# the driver's own invocation at N GPUs (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --impl reference --gpus $N --steps 20 --warmup 5 ) > gpurun_out/benchdrv_${N}gpu_ref.log 2>&1
echo "ref rc=$? $(grep '^{' gpurun_out/benchdrv_${N}gpu_ref.log | tail -1 | cut -c1-160)"
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/benchdrv_${N}gpu.log 2>&1
echo "ours rc=$? $(grep '^{' gpurun_out/benchdrv_${N}gpu.log | tail -1 | cut -c1-260)"
