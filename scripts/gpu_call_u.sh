#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v34}
timeout 200 python scripts/diag_convergence.py 5000 1000 5 60 > gpurun_out/${TAG}_diag_chain.log 2>&1; tail -24 gpurun_out/${TAG}_diag_chain.log
SMCP_B200_NO_CHAIN=1 timeout 300 python scripts/diag_convergence.py 5000 1000 5 60 > gpurun_out/${TAG}_diag_nochain.log 2>&1; tail -24 gpurun_out/${TAG}_diag_nochain.log
