#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v21}
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "bigtop or e2500" > gpurun_out/${TAG}_pytest_big.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_big.log
timeout 600 python scripts/run_config.py C3 3 2000 10000 > gpurun_out/${TAG}_C3.log 2>&1; tail -22 gpurun_out/${TAG}_C3.log
timeout 300 python scripts/run_config.py C5 3 5000 > gpurun_out/${TAG}_C5_n5000.log 2>&1; tail -22 gpurun_out/${TAG}_C5_n5000.log
