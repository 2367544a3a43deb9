#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v33}
timeout 200 python scripts/run_config.py C2 full > gpurun_out/${TAG}_C2_full.log 2>&1; tail -3 gpurun_out/${TAG}_C2_full.log
timeout 200 python scripts/run_config.py C4 full > gpurun_out/${TAG}_C4_full.log 2>&1; tail -3 gpurun_out/${TAG}_C4_full.log
timeout 300 python scripts/run_config.py C3 full 2000 10000 > gpurun_out/${TAG}_C3_full.log 2>&1; tail -3 gpurun_out/${TAG}_C3_full.log
timeout 200 python scripts/run_config.py C5 full 5000 > gpurun_out/${TAG}_C5_n5000_full.log 2>&1; tail -3 gpurun_out/${TAG}_C5_n5000_full.log
RUNCFG_NOPROF=1 timeout 420 python scripts/run_config.py C5 2 20000 > gpurun_out/${TAG}_C5_n20000.log 2>&1; tail -22 gpurun_out/${TAG}_C5_n20000.log
