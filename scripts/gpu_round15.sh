#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for grp in test_dense_path test_end_points test_config003; do
  ( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" --maxfail=30 -p no:cacheprovider ) \
      > gpurun_out/pytest_tc_$grp.log 2>&1
  echo "pytest $grp rc=$? : $(tail -1 gpurun_out/pytest_tc_$grp.log)"
done
( timeout 300 python scripts/bench_dense.py ) 2>&1 | tail -5
( APA_TC_DBG=1 timeout 300 python scripts/bench_dense.py 2>&1 | head -1 )
