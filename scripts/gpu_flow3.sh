#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_flow.py tests/test_gpu_golden.py tests/test_gpu_parity.py -x -q -p no:cacheprovider ) > gpurun_out/pytest_flow3.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/pytest_flow3.log)"
grep -E "^E |FAILED|Error" gpurun_out/pytest_flow3.log | head -10
for args in "" "--streams 1" "--streams 1 --form 2" "--streams 2" "--steps 20 --warmup 5"; do
  tag=$(echo "$args" | tr -d ' -')
  ( timeout 300 python bench.py --cpu-seconds 0.5 --e2e-steps 3 $args ) > gpurun_out/bench3_$tag.log 2>&1
  echo "[$args] $(grep '^{' gpurun_out/bench3_$tag.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('us/step=%.2f value=%.0f frac=%.3f lat=%s launches=%s'%(d['ms_per_step']*1e3,d['value'],d['roofline']['frac'],d.get('latency_us'),d.get('gpu_launches')))")"
done
( timeout 120 python scripts/flow_timeline.py 1 32 ) > gpurun_out/flow_timeline_1lane.txt 2>&1; head -4 gpurun_out/flow_timeline_1lane.txt | cut -c1-330
