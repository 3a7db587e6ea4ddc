#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"apa|gemm" -c 40 --csv --log-file gpurun_out/dense_launches.csv python scripts/bench_dense.py ) > gpurun_out/dense_ncu.log 2>&1
echo rc=$?
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/dense_launches.csv') if not l.startswith('==')]
rows=list(csv.reader(lines)); hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:40]:
    print(r[ki][:70].replace('void apa::<unnamed>::',''), r[gi], '%.1f us'%(float(r[vi])/1e3))
PY
