mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=8 -m gpu 2>&1 | tail -15) > gpurun_out/r02_v21_pytest_kernels.log
(timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -s --maxfail=4 -k "C3 or C5" 2>&1 | tail -12) > gpurun_out/r02_v21_pytest_sizes.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 6 2>&1 | tail -16) > gpurun_out/r02_v21_C3.log
(SMCP_B200_BIG_COMPL_FLOPS=6e5 RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 6 2>&1 | tail -16) > gpurun_out/r02_v21_C3_c6e5.log
tail -n 6 gpurun_out/r02_v21_pytest_kernels.log; cat gpurun_out/r02_v21_pytest_sizes.log; cat gpurun_out/r02_v21_C3.log gpurun_out/r02_v21_C3_c6e5.log
