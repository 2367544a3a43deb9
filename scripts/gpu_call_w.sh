#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v36}
timeout 100 python scripts/diag_esd.py mtx 30 30 40 > gpurun_out/${TAG}_esd_mtx30.log 2>&1; grep -E "status|  it" gpurun_out/${TAG}_esd_mtx30.log | tail -5
timeout 200 python scripts/diag_esd.py mtx 200 200 500 > gpurun_out/${TAG}_esd_C4.log 2>&1; grep -E "status|  it" gpurun_out/${TAG}_esd_C4.log | tail -16
SMCP_B200_BIG_FLOPS=0 timeout 300 python scripts/diag_esd.py mtx 200 200 500 > gpurun_out/${TAG}_esd_C4_nobig.log 2>&1; grep -E "status" gpurun_out/${TAG}_esd_C4_nobig.log | tail -3
