mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=8 -m gpu -x 2>&1 | tail -4) > gpurun_out/r02_v34_pytest_kernels.log
(RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [3568]|op_|kkt_|status") > gpurun_out/r02_v34_C3.log
(timeout 600 python scripts/op_profile.py C3 hessian hessian_inv cholesky 2>&1 | tail -40) > gpurun_out/r02_v34_op_profile_C3.log
tail -n 3 gpurun_out/r02_v34_pytest_kernels.log; cat gpurun_out/r02_v34_C3.log gpurun_out/r02_v34_op_profile_C3.log
