mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=8 -m gpu -x 2>&1 | tail -4) > gpurun_out/r02_v36_pytest_kernels.log
(timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q --maxfail=4 -k "C3" 2>&1 | tail -3) > gpurun_out/r02_v36_pytest_sizes.log
(RUNCFG_NOPROF=1 RUNCFG_PER_ITER=1 timeout 300 python scripts/run_config.py C3 12 2>&1 | grep -E "iteration|op_|kkt_|status") > gpurun_out/r02_v36_C3.log
(timeout 600 python scripts/op_profile.py C3 hessian cholesky 2>&1 | tail -20) > gpurun_out/r02_v36_op_profile_C3.log
tail -n 3 gpurun_out/r02_v36_pytest_kernels.log gpurun_out/r02_v36_pytest_sizes.log; cat gpurun_out/r02_v36_C3.log gpurun_out/r02_v36_op_profile_C3.log
