#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 120 python scripts/timeline.py 32 0 1 ) > gpurun_out/timeline_b32.log 2>&1; echo "timeline rc=$?"
( timeout 120 python scripts/timeline.py 32 1 1 ) > gpurun_out/timeline_b32_c1.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:apa_pool_fused -s 20 -c 2 \
   -o gpurun_out/prof_fused python bench.py --steps 40 --warmup 5 --no-graph --cpu-seconds 0.2 --e2e-steps 2 ) > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
