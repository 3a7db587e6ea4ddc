#!/bin/bash
# A/B of ApaTuning.tail_smem_kb (how many tail CTAs may share an SM) by lanes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lanes in "1 0" "2 0" "6 -1" "4 49"; do
  set -- $lanes
  for pad in 0 40 60 76 115; do
    echo -n "lanes=$1 stream_ctas=$2 tail_smem_kb=$pad: "
    timeout 100 python bench.py --streams $1 --stream-ctas $2 --tail-smem-kb $pad --cpu-seconds 0.3 --e2e-steps 2 --no-latency 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f us/step  (roofline loop %.2f)' % (d['ms_per_step']*1e3, d['roofline']['us_per_step']))"
  done
done 2>&1 | tee gpurun_out/pad_ab.txt
