#!/bin/bash
# A/B of launch structures on the bench workload: args are passed to bench.py, one config per line of $1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/ab_$(basename "$1").log
: > $out
while IFS= read -r cfg; do
  [ -z "$cfg" ] && continue
  echo -n "$cfg: " >> $out
  envp="${cfg%%|*}"; args="${cfg#*|}"
  [ "$envp" = "$cfg" ] && envp=""
  env $envp timeout 300 python bench.py $args --steps 4200 --warmup 84 --cpu-seconds 0.2 --e2e-steps 3 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    j = json.loads(l); print(round(j['ms_per_step']*1000,3), round(j['value']), round(j['roofline']['frac'],3), j['gpu_launches'], j['config'].get('plan'))
except Exception as e:
    print('ERR', l[-300:])
" >> $out
done < "$1"
cat $out
