#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 120 python scripts/timeline.py 32 0 1 50 ) > gpurun_out/tl5_auto.log 2>&1; echo "timeline rc=$?"
( timeout 120 python scripts/timeline.py 32 4 1 50 ) > gpurun_out/tl5_c4.log 2>&1; echo "timeline rc=$?"
( timeout 120 python scripts/timeline.py 32 4 1 100 ) > gpurun_out/tl5_c4_t100.log 2>&1; echo "timeline rc=$?"
( timeout 120 python scripts/timeline.py 32 4 1 -1 ) > gpurun_out/tl5_c4_t0.log 2>&1; echo "timeline rc=$?"
for args in "--streams 2 --cluster 4" "--streams 1 --cluster 4" "--streams 2 --cluster 4 --fused 0" "--streams 2 --cluster 2" "--streams 2 --cluster 16"; do
  tag=$(echo "$args" | tr -d ' ')
  ( timeout 300 python bench.py --steps 2002 --warmup 20 --cpu-seconds 1 --e2e-steps 3 $args ) > gpurun_out/bench5_$tag.log 2>&1
  echo "bench [$args] rc=$? : $(tail -1 gpurun_out/bench5_$tag.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f plan=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['config']['plan']))
except Exception as e: print('parse fail', e)
")"
done
