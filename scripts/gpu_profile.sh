#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one full capture of the headline kernel.
# (--no-pdl --no-graph --streams 1: the profiler serialises kernels; nothing to overlap.)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 80 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 30 --warmup 3 --no-graph --no-pdl --streams 1 --cpu-seconds 0.2 --e2e-steps 2 ) \
    > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:apa_pool -s 12 -c 2 \
    -o gpurun_out/prof_fused python bench.py --steps 30 --warmup 3 --no-graph --no-pdl --streams 1 --cpu-seconds 0.2 --e2e-steps 2 ) \
    > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
( timeout 200 python scripts/timeline.py 32 0 1 1 ) > gpurun_out/timeline_local_solo.log 2>&1
( timeout 200 python scripts/timeline.py 32 0 1 0 40 ) > gpurun_out/timeline_local_steady.log 2>&1
( timeout 300 python bench.py --steps 2002 --warmup 20 ) > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
( timeout 300 python bench.py --steps 2002 --warmup 20 --streams 1 ) > gpurun_out/bench_streams1.log 2>&1
( timeout 300 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"
( timeout 300 python scripts/bench_dense.py ) > gpurun_out/bench_dense.log 2>&1
