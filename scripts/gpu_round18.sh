#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for grp in test_cluster_local test_fused_path_parity test_back_to_back test_bitwise test_two_batches test_config003 test_properties test_host_buffer test_network_fn; do
  ( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" --maxfail=30 -p no:cacheprovider ) \
      > gpurun_out/pytest_l_$grp.log 2>&1
  echo "pytest $grp rc=$? : $(tail -1 gpurun_out/pytest_l_$grp.log)"
done
for args in "" "--general" "--no-pdl" "--streams 2" "--no-graph"; do
  tag=$(echo "$args" | tr -d ' ')
  ( timeout 300 python bench.py --steps 2002 --warmup 20 --cpu-seconds 1 --e2e-steps 3 $args ) > gpurun_out/bench18_$tag.log 2>&1
  echo "bench [$args] rc=$? : $(tail -1 gpurun_out/bench18_$tag.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f launch=%s plan=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d['config']['launch'],d['config']['plan']))
except Exception as e: print('parse fail', e)
")"
done
grep -E "Error|error|FAILED|assert" gpurun_out/pytest_l_test_cluster_local.log | head -10
