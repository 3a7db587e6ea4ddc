#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tp in 50 100 -1 33; do
  ( timeout 120 python scripts/timeline.py 32 0 1 $tp ) > gpurun_out/tl3_t$tp.log 2>&1; echo "timeline tma=$tp rc=$?"
done
for grp in test_fused_path_parity test_cluster_sizes test_load_path test_two_batches test_config003 test_end_points test_bitwise test_properties test_network_fn test_host_buffer; do
  ( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" --maxfail=30 -x -p no:cacheprovider ) \
      > gpurun_out/pytest_$grp.log 2>&1
  echo "pytest $grp rc=$? : $(tail -1 gpurun_out/pytest_$grp.log)"
done
for args in "--streams 2" "--streams 1" "--streams 2 --tma-percent 100" "--streams 2 --tma-percent -1" "--streams 2 --tma-percent 33" "--streams 2 --tma-percent 67" "--streams 1 --tma-percent -1"; do
  tag=$(echo "$args" | tr -d ' ')
  ( timeout 300 python bench.py --steps 2002 --warmup 20 --cpu-seconds 2 --e2e-steps 5 $args ) > gpurun_out/bench4_$tag.log 2>&1
  echo "bench [$args] rc=$? : $(tail -1 gpurun_out/bench4_$tag.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f cpu=%.0f@%d'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['cpu_baseline']['value'],d['cpu_baseline']['cores']))
except Exception as e: print('parse fail', e)
")"
done
