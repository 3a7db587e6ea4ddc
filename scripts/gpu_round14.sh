#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 0 1 2 3; do
( APA_TC_DBG=$d timeout 300 python scripts/bench_dense.py 2>&1 | head -1 ) ; 
done
