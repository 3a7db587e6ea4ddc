#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_fused_gather or test_cluster_local or test_back_to_back" --maxfail=30 -p no:cacheprovider ) > gpurun_out/pytest_gather.log 2>&1
echo "pytest gather rc=$? : $(tail -1 gpurun_out/pytest_gather.log)"
grep -E "Error|FAILED|assert " gpurun_out/pytest_gather.log | head
