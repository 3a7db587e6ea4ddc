#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 120 python scripts/timeline.py 32 4 1 50 0 ) > gpurun_out/tl6_cold.log 2>&1; echo "rc=$?"
( timeout 120 python scripts/timeline.py 32 4 1 50 1 ) > gpurun_out/tl6_warm.log 2>&1; echo "rc=$?"
( timeout 120 python scripts/timeline.py 32 4 1 100 1 ) > gpurun_out/tl6_warm_t100.log 2>&1; echo "rc=$?"
( timeout 120 python scripts/timeline.py 32 4 1 -1 1 ) > gpurun_out/tl6_warm_t0.log 2>&1; echo "rc=$?"
