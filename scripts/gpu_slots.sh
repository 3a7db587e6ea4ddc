#!/bin/bash
# A/B of the cluster-local kernel's shapes x batches in flight, plus the under-load timeline of the widest.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster_local or back_to_back or two_batches or gather_single" > gpurun_out/slots_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/slots_pytest.log
tail -5 gpurun_out/slots_pytest.log
: > gpurun_out/slots_bench.log
for cfg in "2 1" "2 2" "2 3" "2 4" "1 3"; do
  set -- $cfg
  echo -n "shape $1 streams $2: " >> gpurun_out/slots_bench.log
  timeout 300 python bench.py --shape $1 --streams $2 --steps 4200 --warmup 84 --cpu-seconds 0.2 --e2e-steps 3 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    j = json.loads(l); print(round(j['ms_per_step']*1000,3), round(j['value']), round(j['roofline']['frac'],3), j['config'].get('plan'))
except Exception as e:
    print('ERR', l[-300:])
" >> gpurun_out/slots_bench.log
done
cat gpurun_out/slots_bench.log
for a in "3 2" "1 2"; do timeout 100 python scripts/timeline_load.py $a 2>&1 | tail -20; done > gpurun_out/timeline_load2.txt 2>&1
