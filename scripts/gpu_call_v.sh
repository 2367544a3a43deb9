#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v35}
SMCP_B200_CHAIN_DEBUG=1 SMCP_B200_CHAIN_GAMMA_MAX=1e300 timeout 200 python scripts/diag_convergence.py 5000 1000 5 45 > gpurun_out/${TAG}_diag_gamma.log 2>&1
grep "max |Phi|" gpurun_out/${TAG}_diag_gamma.log | awk '{print $7}' | tr '\n' ' ' | head -c 1500; echo
grep "status\|  it" gpurun_out/${TAG}_diag_gamma.log | head -16
for G in 1e1 1e3 1e5 1e7; do
  SMCP_B200_CHAIN_GAMMA_MAX=$G timeout 200 python scripts/diag_convergence.py 5000 1000 5 60 > gpurun_out/${TAG}_diag_g$G.log 2>&1
  grep "status" gpurun_out/${TAG}_diag_g$G.log
done
