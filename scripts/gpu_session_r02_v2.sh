mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -q --maxfail=8 2>&1 | tail -25) > gpurun_out/r02_v2_pytest_dense.log
(SMCP_B200_PT_DEBUG=1 timeout 300 python scripts/bench_kernels.py potrf 2>&1 | tail -60) > gpurun_out/r02_v2_potrf_phases.log
(timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_dense.py 2>&1 | tail -40) > gpurun_out/r02_v2_pytest_gpu.log
(timeout 400 python scripts/run_config.py C3 4 2>&1 | tail -30) > gpurun_out/r02_v2_C3.log
tail -n 8 gpurun_out/r02_v2_pytest_dense.log; tail -n 12 gpurun_out/r02_v2_pytest_gpu.log; grep -E "potrf_tile m=(1000|1186|2560) |^potrf" gpurun_out/r02_v2_potrf_phases.log | sort | uniq | head -30; tail -n 22 gpurun_out/r02_v2_C3.log
