#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pieces in 1 4 16 32; do
  ( timeout 120 python scripts/timeline.py 32 0 1 $pieces ) > gpurun_out/tl_p$pieces.log 2>&1; echo "timeline pieces=$pieces rc=$?"
done
( timeout 120 python scripts/timeline.py 32 1 1 4 ) > gpurun_out/tl_c1_p4.log 2>&1
( timeout 120 python scripts/timeline.py 32 2 1 4 ) > gpurun_out/tl_c2_p4.log 2>&1
for grp in test_fused_path_parity test_cluster_sizes test_config003 test_end_points test_bitwise test_properties test_network_fn; do
  ( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" --maxfail=30 -x -p no:cacheprovider ) \
      > gpurun_out/pytest_$grp.log 2>&1
  echo "pytest $grp rc=$? : $(tail -1 gpurun_out/pytest_$grp.log)"
done
( timeout 300 python bench.py --steps 2000 --warmup 20 --cpu-seconds 2 ) > gpurun_out/bench_r3.log 2>&1
echo "bench rc=$? : $(tail -1 gpurun_out/bench_r3.log | cut -c1-300)"
