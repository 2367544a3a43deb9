mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -q --maxfail=8 -m gpu 2>&1 | tail -15) > gpurun_out/r02_v11_pytest_dense.log
(timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_dense.py --deselect tests/test_gpu_baseline_sizes.py 2>&1 | tail -30) > gpurun_out/r02_v11_pytest_gpu.log
(timeout 300 python scripts/bench_kernels.py trsm 2>&1 | tail -16) > gpurun_out/r02_v11_bench_trsm.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -26) > gpurun_out/r02_v11_C3_9it.log
tail -n 6 gpurun_out/r02_v11_pytest_dense.log; tail -n 8 gpurun_out/r02_v11_pytest_gpu.log; cat gpurun_out/r02_v11_bench_trsm.log; cat gpurun_out/r02_v11_C3_9it.log
