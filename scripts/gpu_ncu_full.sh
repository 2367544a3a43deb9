#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v29}
for K in gemm_dmma_kernel chain_staged_batch_kernel chain_staged_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 4 \
      -f -o gpurun_out/${TAG}_$K python scripts/profile_iter.py 5000 1000 5 2 > gpurun_out/${TAG}_ncu_$K.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${TAG}_$K.ncu-rep > gpurun_out/${TAG}_ncu_$K.txt 2>&1
  grep -E "kernel:|gpu__time_duration|dram__bytes|grid_size|dmma_cycles|dram_throughput" gpurun_out/${TAG}_ncu_$K.txt | head -28
done
