#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_eval_tail.py tests/test_gpu_backward.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_evaltail.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/pytest_evaltail.log)"
grep -E "^E |FAILED|Error" gpurun_out/pytest_evaltail.log | head -20
bash scripts/gpu_sanitize.sh all memcheck
bash scripts/gpu_sanitize.sh all racecheck
bash scripts/gpu_sanitize.sh all synccheck
bash scripts/gpu_sanitize.sh train initcheck
