mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=8 -m gpu -x 2>&1 | tail -6) > gpurun_out/r02_v28_pytest_kernels.log
(timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -s --maxfail=4 -k "C3" 2>&1 | tail -5) > gpurun_out/r02_v28_pytest_sizes.log
(RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [35678]|op_|kkt_|status") > gpurun_out/r02_v28_C3.log
(timeout 600 python scripts/op_profile.py C3 completion hessian_prep_inv 2>&1 | tail -30) > gpurun_out/r02_v28_op_profile_C3.log
tail -n 3 gpurun_out/r02_v28_pytest_kernels.log; cat gpurun_out/r02_v28_pytest_sizes.log gpurun_out/r02_v28_C3.log gpurun_out/r02_v28_op_profile_C3.log
