#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for args in "--no-coop" "" "--no-coop --fused 0" "--no-coop --streams 2" "--no-coop --early-trigger 1"; do
  tag=$(echo "$args" | tr -d ' ')
  ( timeout 300 python bench.py --steps 2002 --warmup 20 --cpu-seconds 1 --e2e-steps 3 $args ) > gpurun_out/bench12_$tag.log 2>&1
  echo "bench [$args] rc=$? : $(tail -1 gpurun_out/bench12_$tag.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f launch=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['config']['launch']))
except Exception as e: print('parse fail', e)
")"
done
