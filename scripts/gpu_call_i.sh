#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v22}
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "bigtop or e2500" > gpurun_out/${TAG}_pytest_big.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_big.log
timeout 600 python -m pytest tests/test_gpu_solver.py -x -q > gpurun_out/${TAG}_pytest_solver.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_solver.log
timeout 300 python scripts/run_config.py C4 4 > gpurun_out/${TAG}_C4.log 2>&1; tail -22 gpurun_out/${TAG}_C4.log
