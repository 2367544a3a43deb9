#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v17}
for LA in 1 0; do
  echo "== lookahead=$LA"
  SMCP_B200_LOOKAHEAD=$LA timeout 300 python scripts/host_prof.py 2>&1 | tee gpurun_out/${TAG}_hostprof_la$LA.log | tail -18
done
