#!/bin/bash
# What the driver runs at round end, in one go: the whole GPU suite in ONE process, smoke(), the bench (all configs).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest -m gpu rc=$? : $(tail -1 gpurun_out/pytest_gpu_full.log)"
grep -E "FAILED|Error" gpurun_out/pytest_gpu_full.log | head -5
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -2 gpurun_out/smoke.log | tr '\n' ' ')"
( timeout 300 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_driver_window.log 2>&1; echo "bench(20/5) rc=$?"; tail -1 gpurun_out/bench_driver_window.log | cut -c1-200
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference.log | cut -c1-200
( timeout 300 python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-200
( timeout 300 python bench.py --streams 1 --no-latency --cpu-seconds 0.5 --e2e-steps 3 ) > gpurun_out/bench_streams1.log 2>&1; tail -1 gpurun_out/bench_streams1.log | cut -c1-120
for c in 2 3 4 5; do
  ( timeout 300 python bench.py --config $c --cpu-seconds 2 ) > gpurun_out/bench_config$c.log 2>&1; echo "config $c rc=$?"; tail -1 gpurun_out/bench_config$c.log | cut -c1-200
done
( timeout 300 python scripts/bench_dense.py ) > gpurun_out/bench_dense.log 2>&1; tail -6 gpurun_out/bench_dense.log
