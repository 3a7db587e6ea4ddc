#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 120 python scripts/timeline.py 32 4 1 0 0 ) > gpurun_out/tl9_c4.log 2>&1; echo "rc=$?"
( timeout 120 python scripts/timeline.py 32 0 1 0 0 ) > gpurun_out/tl9_auto.log 2>&1; echo "rc=$?"
for grp in test_fused_path_parity test_cluster_sizes test_load_path test_two_batches test_config003 test_end_points test_bitwise test_properties; do
  ( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" --maxfail=30 -x -p no:cacheprovider ) \
      > gpurun_out/pytest_$grp.log 2>&1
  echo "pytest $grp rc=$? : $(tail -1 gpurun_out/pytest_$grp.log)"
done
for args in "--streams 2 --cluster 4" "--streams 1 --cluster 4" "--streams 2" "--streams 1" "--streams 2 --cluster 2" "--streams 2 --cluster 4 --fused 0"; do
  tag=$(echo "$args" | tr -d ' ')
  ( timeout 300 python bench.py --steps 2002 --warmup 20 --cpu-seconds 1 --e2e-steps 3 $args ) > gpurun_out/bench9_$tag.log 2>&1
  echo "bench [$args] rc=$? : $(tail -1 gpurun_out/bench9_$tag.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f plan=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['config']['plan']))
except Exception as e: print('parse fail', e)
")"
done
