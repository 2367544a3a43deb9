mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -q --maxfail=8 -m gpu 2>&1 | tail -15) > gpurun_out/r02_v19_pytest_dense.log
(SMCP_B200_PT_DEBUG=1 timeout 300 python scripts/bench_kernels.py potrf 2>&1 | grep -E "^potrf|m=(1000|1186|1131|2000|2560) ") > gpurun_out/r02_v19_potrf_phases.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 6 2>&1 | tail -22) > gpurun_out/r02_v19_C3.log
tail -n 5 gpurun_out/r02_v19_pytest_dense.log; awk '!seen[$2 $3]++' gpurun_out/r02_v19_potrf_phases.log | head -24; cat gpurun_out/r02_v19_C3.log
