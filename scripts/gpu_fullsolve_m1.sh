#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v38}
RUNCFG_METHOD=feas timeout 150 python scripts/run_config.py C5 full 5000 > gpurun_out/${TAG}_C5_n5000_feas_full.log 2>&1; tail -4 gpurun_out/${TAG}_C5_n5000_feas_full.log
RUNCFG_METHOD=feas timeout 200 python scripts/run_config.py C4 full > gpurun_out/${TAG}_C4_feas_full.log 2>&1; tail -5 gpurun_out/${TAG}_C4_feas_full.log
