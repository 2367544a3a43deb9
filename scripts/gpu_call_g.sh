#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v20}
timeout 600 python scripts/run_config.py C3 3 2000 10000 > gpurun_out/${TAG}_C3.log 2>&1; tail -24 gpurun_out/${TAG}_C3.log
timeout 300 python scripts/run_config.py C5 3 5000 > gpurun_out/${TAG}_C5_n5000.log 2>&1; tail -22 gpurun_out/${TAG}_C5_n5000.log
timeout 300 python scripts/run_config.py C4 4 > gpurun_out/${TAG}_C4.log 2>&1; tail -22 gpurun_out/${TAG}_C4.log
