#!/bin/bash
# compute-sanitizer over small shapes of every kernel family (scripts/sanitize_cases.py); summaries -> gpurun_out/
# usage: gpu_sanitize.sh [flow|cluster|dense|train|all] [tools...]      (default: all memcheck racecheck synccheck)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
which=${1:-all}; shift
tools=${@:-memcheck racecheck synccheck}
for tool in $tools; do
  log=gpurun_out/sanitizer_${tool}_${which}.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python scripts/sanitize_cases.py $which > $log 2>&1
  echo "exit $?" >> $log
  echo "== $tool ($which): $(grep -c 'rel err' $log) cases, $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)  [$(tail -1 $log)]"
done
