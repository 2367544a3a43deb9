mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_golden.py tests/test_host_symbolic.py -q -m gpu 2>&1 | tail -5) > gpurun_out/r02_v12_pytest_golden.log
(timeout 600 python scripts/op_profile.py C3 completion hessian hessian_inv cholesky 2>&1 | tail -60) > gpurun_out/r02_v12_op_profile_C3.log
(RUNCFG_METHOD=feas RUNCFG_NOPROF=1 timeout 1500 python scripts/run_config.py C5 full 20000 2>&1 | tail -24) > gpurun_out/r02_v12_C5_n20000_feas.log
tail -n 4 gpurun_out/r02_v12_pytest_golden.log; cat gpurun_out/r02_v12_op_profile_C3.log; cat gpurun_out/r02_v12_C5_n20000_feas.log
