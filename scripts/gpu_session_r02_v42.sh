mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=8 -m gpu -x -k "trsm or schur or operator" 2>&1 | tail -4) > gpurun_out/r02_v42_pytest_kernels.log
(timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_golden.py -q --maxfail=4 -k "C3 or C5 or cuda" 2>&1 | tail -3) > gpurun_out/r02_v42_pytest_sizes.log
(RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [3568]|op_trsm|kkt_assemble|status") > gpurun_out/r02_v42_C3.log
tail -n 3 gpurun_out/r02_v42_pytest_kernels.log gpurun_out/r02_v42_pytest_sizes.log; cat gpurun_out/r02_v42_C3.log
