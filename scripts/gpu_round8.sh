#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 8 12 15; do
( timeout 120 python scripts/timeline.py 32 4 1 50 0 $dbg ) > gpurun_out/tl8_d$dbg.log 2>&1; echo "dbg=$dbg rc=$? $(grep phase1_end gpurun_out/tl8_d$dbg.log)"
done
