#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_backward.py -x -q -p no:cacheprovider ) > gpurun_out/pytest_backward.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/pytest_backward.log)"
grep -E "^E |FAILED|Error" gpurun_out/pytest_backward.log | head -20
( timeout 200 python scripts/bench_train.py 32 0.2 ) > gpurun_out/bench_train.log 2>&1; tail -1 gpurun_out/bench_train.log
( timeout 200 python scripts/bench_train.py 32 1.0 ) > gpurun_out/bench_train_keep1.log 2>&1; tail -1 gpurun_out/bench_train_keep1.log
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 22 -c 33 --csv --log-file gpurun_out/train_launches.csv python scripts/bench_train.py 32 0.2 profile ) > /dev/null 2>&1
grep -v "^==" gpurun_out/train_launches.csv | python -c "
import csv,sys,collections
rows=list(csv.reader(sys.stdin)); h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    n=r[ki].split('(')[0][-26:]; a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[vi])/1e3
for n,(c,t) in agg.items(): print('%-28s n=%d avg %8.2f us'%(n,c,t/c))
"
