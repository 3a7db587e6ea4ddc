#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one full capture of the fused kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 80 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 30 --warmup 3 --no-graph --no-pdl --no-coop --cpu-seconds 0.2 --e2e-steps 2 ) \
    > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:apa_pool_fused -s 12 -c 2 \
    -o gpurun_out/prof_fused python bench.py --steps 30 --warmup 3 --no-graph --no-pdl --no-coop --cpu-seconds 0.2 --e2e-steps 2 ) \
    > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
( timeout 300 python bench.py --steps 2002 --warmup 20 ) > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
