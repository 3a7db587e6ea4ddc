#!/bin/bash
# First GPU pass: smoke, parity tests (one process per group so a trap cannot poison the rest), bench variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -5 gpurun_out/smoke.log
for grp in test_fused_path_parity test_cluster_sizes test_config003 test_dense_path test_end_points \
           test_bitwise test_properties test_host_buffer test_network_fn test_gram test_errors; do
  ( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" --maxfail=30 -x -p no:cacheprovider ) \
      > gpurun_out/pytest_$grp.log 2>&1
  echo "pytest $grp rc=$? : $(tail -1 gpurun_out/pytest_$grp.log)"
done
for args in "" "--fused 0" "--cluster 2" "--cluster 1" "--cluster 8" "--no-graph"; do
  tag=$(echo "$args" | tr -d ' -')
  ( timeout 300 python bench.py --steps 2000 --warmup 20 --cpu-seconds 3 $args ) > gpurun_out/bench_$tag.log 2>&1
  echo "bench [$args] rc=$? : $(tail -1 gpurun_out/bench_$tag.log | cut -c1-600)"
done
