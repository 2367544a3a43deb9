mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_golden.py -q --maxfail=8 -m gpu 2>&1 | tail -25) > gpurun_out/r02_v3_pytest_dense.log
(SMCP_B200_PT_DEBUG=1 timeout 300 python scripts/bench_kernels.py potrf 2>&1 | grep -E "^potrf|m=(1000|1186|1131|2000|2560) ") > gpurun_out/r02_v3_potrf_phases.log
(timeout 300 python scripts/bench_kernels.py trsm gemm 2>&1 | tail -40) > gpurun_out/r02_v3_bench_kernels.log
(timeout 400 python scripts/run_config.py C3 4 2>&1 | tail -30) > gpurun_out/r02_v3_C3.log
(timeout 900 python bench.py 2>gpurun_out/r02_v3_bench.err | tail -1) > gpurun_out/r02_v3_bench.json
tail -n 6 gpurun_out/r02_v3_pytest_dense.log; awk '!seen[$2 $3]++' gpurun_out/r02_v3_potrf_phases.log | head -30; cat gpurun_out/r02_v3_bench_kernels.log; tail -n 24 gpurun_out/r02_v3_C3.log; tail -n 5 gpurun_out/r02_v3_bench.err; head -c 3000 gpurun_out/r02_v3_bench.json
