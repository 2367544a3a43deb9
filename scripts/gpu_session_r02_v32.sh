mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dense.py -q --maxfail=8 -m gpu -x 2>&1 | tail -4) > gpurun_out/r02_v32_pytest_kernels.log
(timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_solver.py -q --maxfail=4 -k "C3 or dense_potrf or feas" 2>&1 | tail -4) > gpurun_out/r02_v32_pytest_sizes.log
(RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [35678]|op_|kkt_|status") > gpurun_out/r02_v32_C3.log
(SMCP_B200_NO_SIDE=1 RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 6 2>&1 | grep -E "iteration [35]|op_hessian ") > gpurun_out/r02_v32_C3_noside.log
tail -n 3 gpurun_out/r02_v32_pytest_kernels.log gpurun_out/r02_v32_pytest_sizes.log; cat gpurun_out/r02_v32_C3.log gpurun_out/r02_v32_C3_noside.log
