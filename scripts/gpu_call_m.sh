#!/bin/bash
# ncu --set full captures (with source) of the latency kernels of the bench workload
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v26}
for K in potrs_kernel potrf_diag_kernel chain_chol_staged_kernel chain_scan2_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 3 \
      -f -o gpurun_out/${TAG}_$K python scripts/profile_iter.py 5000 1000 5 7 > gpurun_out/${TAG}_ncu_$K.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${TAG}_$K.ncu-rep > gpurun_out/${TAG}_ncu_$K.txt 2>&1
  grep -E "kernel:|gpu__time_duration|grid_size|block_size" gpurun_out/${TAG}_ncu_$K.txt | head -12
done
ls -la gpurun_out/*.ncu-rep
