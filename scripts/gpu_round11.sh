#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for args in "" "--early-trigger 1" "--fused 0" "--fused 0 --early-trigger 1" "--streams 2" "--streams 2 --early-trigger 1"; do
  tag=$(echo "$args" | tr -d ' ')
  ( timeout 300 python bench.py --steps 2002 --warmup 20 --cpu-seconds 1 --e2e-steps 3 $args ) > gpurun_out/bench11_$tag.log 2>&1
  echo "bench [$args] rc=$? : $(tail -1 gpurun_out/bench11_$tag.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f launch=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['config']['launch']))
except Exception as e: print('parse fail', e)
")"
done
