#!/bin/bash
# parity + A/B + under-load timeline of the two-kernel STREAM + CONTRACT pipeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or gather_single" > gpurun_out/split_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/split_pytest.log
tail -15 gpurun_out/split_pytest.log
cat > /tmp/ab_split.txt <<'EOT'
--split --streams 1
--split --streams 2
--split --streams 3
--split --streams 4
--split --streams 2 --cluster 8
--split --streams 1 --no-pdl
EOT
bash scripts/gpu_ab.sh /tmp/ab_split.txt
for a in "1 7" "2 7"; do timeout 100 python scripts/timeline_load.py $a 2>&1 | tail -20; done > gpurun_out/timeline_split.txt 2>&1
