#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v19}
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/${TAG}_pytest_kernels.log 2>&1
tail -25 gpurun_out/${TAG}_pytest_kernels.log
timeout 600 python -m pytest tests/test_gpu_solver.py tests/test_gpu_chain.py -x -q > gpurun_out/${TAG}_pytest_solver.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_solver.log
timeout 600 python scripts/run_config.py C3 3 2000 10000 > gpurun_out/${TAG}_C3.log 2>&1; tail -22 gpurun_out/${TAG}_C3.log
timeout 300 python scripts/run_config.py C5 3 5000 > gpurun_out/${TAG}_C5_n5000.log 2>&1; tail -20 gpurun_out/${TAG}_C5_n5000.log
timeout 300 python scripts/run_config.py C4 4 > gpurun_out/${TAG}_C4.log 2>&1; tail -20 gpurun_out/${TAG}_C4.log
