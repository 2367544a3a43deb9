mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -q --maxfail=8 -m gpu 2>&1 | tail -6) > gpurun_out/r02_v30_pytest_dense.log
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=8 -m gpu -x 2>&1 | tail -6) > gpurun_out/r02_v30_pytest_kernels.log
(SMCP_B200_PT_DEBUG=1 timeout 300 python scripts/bench_kernels.py potrf 2>&1 | grep -E "^potrf|m=(1000|1186|2560) ") > gpurun_out/r02_v30_potrf_phases.log
(timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -s --maxfail=4 -k "C3" 2>&1 | tail -5) > gpurun_out/r02_v30_pytest_sizes.log
(RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [35678]|op_|kkt_|status") > gpurun_out/r02_v30_C3.log
tail -n 3 gpurun_out/r02_v30_pytest_dense.log gpurun_out/r02_v30_pytest_kernels.log; awk '!seen[$2 $3]++' gpurun_out/r02_v30_potrf_phases.log | head -20; cat gpurun_out/r02_v30_pytest_sizes.log gpurun_out/r02_v30_C3.log
