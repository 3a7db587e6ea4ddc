#!/bin/bash
# What the driver runs at round end, in one go: the whole GPU suite in ONE process, smoke(), the bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest -m gpu rc=$? : $(tail -1 gpurun_out/pytest_gpu_full.log)"
grep -E "FAILED|Error" gpurun_out/pytest_gpu_full.log | head -5
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -2 gpurun_out/smoke.log | tr '\n' ' ')"
( timeout 300 python scripts/bench_dense.py ) > gpurun_out/bench_dense.log 2>&1; cat gpurun_out/bench_dense.log | tail -6
( timeout 300 python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-160
