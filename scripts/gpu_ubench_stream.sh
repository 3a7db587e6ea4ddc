#!/bin/bash
# round-2 first experiment: what limits short overlapped streams.  ~1 GPU-minute.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench_stream.bin scripts/ubench_stream.cu || exit 1
(timeout 200 ./scripts/ubench_stream.bin zeros; timeout 200 ./scripts/ubench_stream.bin) | tee gpurun_out/ubench_stream_data.txt
