mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=8 -m gpu -x 2>&1 | tail -8) > gpurun_out/r02_v26_pytest_kernels.log
(timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -s --maxfail=4 -k "C3" 2>&1 | tail -6) > gpurun_out/r02_v26_pytest_sizes.log
(RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 6 2>&1 | grep -E "iteration [35]|op_|kkt_") > gpurun_out/r02_v26_C3.log
(SMCP_B200_NO_DINV=1 RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 6 2>&1 | grep -E "iteration [35]|op_hessian ") > gpurun_out/r02_v26_C3_nodinv.log
(timeout 600 python scripts/op_profile.py C3 hessian hessian_inv completion 2>&1 | tail -50) > gpurun_out/r02_v26_op_profile_C3.log
tail -n 4 gpurun_out/r02_v26_pytest_kernels.log; cat gpurun_out/r02_v26_pytest_sizes.log; cat gpurun_out/r02_v26_C3.log gpurun_out/r02_v26_C3_nodinv.log; cat gpurun_out/r02_v26_op_profile_C3.log
