#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for args in "--trigger-lead -1" "--trigger-lead 2" "--trigger-lead 4" "--trigger-lead 6" "--trigger-lead 9" "--trigger-lead 12" "--general --trigger-lead -1" "--general --trigger-lead 4" "--general --trigger-lead 8" "--general --trigger-lead 12" "--streams 2 --trigger-lead 6" "--general --streams 2 --trigger-lead 6"; do
  tag=$(echo "$args" | tr -d ' ')
  ( timeout 300 python bench.py --steps 2002 --warmup 20 --cpu-seconds 0.5 --e2e-steps 2 $args ) > gpurun_out/bench20_$tag.log 2>&1
  echo "bench [$args] rc=$? : $(tail -1 gpurun_out/bench20_$tag.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac']))
except Exception as e: print('parse fail', e)
")"
done
