#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_eval_tail.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_evaltail.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/pytest_evaltail.log)"
for fam in train dense cluster; do bash scripts/gpu_sanitize.sh $fam racecheck; done
