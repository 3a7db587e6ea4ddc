#!/bin/bash
# PDL trigger-lead sweep for the 3-slot shape of the cluster-local kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/slots_lead.log
for st in 1 2 3 4; do
  for lead in -1 2 5 8 11 14 0; do
    echo -n "slots 3 streams $st lead $lead: " >> gpurun_out/slots_lead.log
    timeout 300 python bench.py --slots 3 --streams $st --trigger-lead $lead --steps 4200 --warmup 84 --cpu-seconds 0.2 --e2e-steps 3 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    j = json.loads(l); print(round(j['ms_per_step']*1000,3), round(j['value']), round(j['roofline']['frac'],3))
except Exception as e:
    print('ERR', l[-300:])
" >> gpurun_out/slots_lead.log
  done
done
cat gpurun_out/slots_lead.log
