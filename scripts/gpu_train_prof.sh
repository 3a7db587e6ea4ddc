#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fused|small|contract" -s 4 -c 4 -f -o gpurun_out/prof_train python scripts/bench_train.py 32 0.2 profile ) > gpurun_out/prof_train.log 2>&1; echo "rc=$?"; ls -la gpurun_out/prof_train.ncu-rep
