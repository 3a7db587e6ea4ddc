#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "--no-pdl --fused 0" "--no-pdl --no-coop" "--no-pdl" "--no-coop"; do
  tag=$(echo "$v" | tr -d ' ')
  ( timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:apa_pool -s 2 -c 6 --csv \
    --log-file gpurun_out/dbg_$tag.csv python bench.py --steps 10 --warmup 3 --no-graph $v --cpu-seconds 0.1 --e2e-steps 1 ) \
    > gpurun_out/dbg_$tag.log 2>&1; echo "[$v] rc=$? $(grep -c apa_pool gpurun_out/dbg_$tag.csv) rows; $(grep -i 'error' gpurun_out/dbg_$tag.log | head -2 | tr '\n' ' ')"
done
