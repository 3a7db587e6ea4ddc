#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_fused_gather or test_cluster_local or test_back_to_back or test_bitwise or test_two_batches or test_fused_path_parity" --maxfail=30 -p no:cacheprovider ) > gpurun_out/pytest_l2.log 2>&1
echo "pytest rc=$? : $(tail -1 gpurun_out/pytest_l2.log)"
grep -E "Error|FAILED|assert " gpurun_out/pytest_l2.log | head
( timeout 100 python scripts/timeline.py 32 0 1 1 ) > gpurun_out/tl23_solo.log 2>&1; sed -n 2,14p gpurun_out/tl23_solo.log
for args in "" "--streams 1"; do
  ( timeout 300 python bench.py --steps 2002 --warmup 20 --cpu-seconds 0.5 --e2e-steps 2 $args ) 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac']))"
done
