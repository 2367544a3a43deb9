mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -q --maxfail=8 -m gpu 2>&1 | tail -25) > gpurun_out/r02_v6_pytest_dense.log
(SMCP_B200_PT_DEBUG=1 timeout 300 python scripts/bench_kernels.py potrf 2>&1 | grep -E "^potrf|m=(1000|1186|1131|2000|2560) ") > gpurun_out/r02_v6_potrf_phases.log
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=10 -k "half" 2>&1 | tail -30) > gpurun_out/r02_v6_pytest_half.log
(timeout 1500 python -m pytest tests/test_gpu_baseline_sizes.py -q -s --maxfail=10 2>&1 | tail -60) > gpurun_out/r02_v6_pytest_sizes.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -26) > gpurun_out/r02_v6_C3_9it.log
tail -n 6 gpurun_out/r02_v6_pytest_dense.log; awk '!seen[$2 $3]++' gpurun_out/r02_v6_potrf_phases.log | head -30; tail -n 12 gpurun_out/r02_v6_pytest_half.log; tail -n 40 gpurun_out/r02_v6_pytest_sizes.log; cat gpurun_out/r02_v6_C3_9it.log
