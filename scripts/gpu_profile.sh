#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one full capture of the headline kernel.
# (--no-pdl --no-graph --streams 1: the profiler serialises kernels; nothing to overlap.)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 3 --no-graph --no-pdl --streams 1 --stream-ctas 0 --cpu-seconds 0.2 --e2e-steps 2 --no-latency"
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 120 --csv \
    --log-file gpurun_out/launches.csv $B ) > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:apa_flow_stream -s 6 -c 2 \
    -o gpurun_out/prof_stream $B ) > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/prof_stream.ncu-rep
