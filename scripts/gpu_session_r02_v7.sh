mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -q --maxfail=8 -m gpu 2>&1 | tail -25) > gpurun_out/r02_v7_pytest_dense.log
(SMCP_B200_PT_DEBUG=1 timeout 300 python scripts/bench_kernels.py potrf 2>&1 | grep -E "^potrf|m=(1000|1186|1131|2000|2560) ") > gpurun_out/r02_v7_potrf_phases.log
(timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -q -s --maxfail=10 -k "C5 or C3" 2>&1 | tail -30) > gpurun_out/r02_v7_pytest_sizes.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -26) > gpurun_out/r02_v7_C3_9it.log
(timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_dense.py --deselect tests/test_gpu_baseline_sizes.py 2>&1 | tail -30) > gpurun_out/r02_v7_pytest_gpu.log
# ncu captures of the dense kernels (one launch each)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_tn -c 1 -o gpurun_out/r02_v7_ncu_gemm_tma python scripts/bench_kernels.py gemm > gpurun_out/r02_v7_ncu_gemm_tma.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_tile -s 2 -c 1 -o gpurun_out/r02_v7_ncu_potrf_tile python scripts/bench_kernels.py potrf > gpurun_out/r02_v7_ncu_potrf_tile.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:trsm_slab -s 2 -c 1 -o gpurun_out/r02_v7_ncu_trsm_slab python scripts/bench_kernels.py trsm > gpurun_out/r02_v7_ncu_trsm_slab.log 2>&1)
ls -la gpurun_out/*.ncu-rep
tail -n 6 gpurun_out/r02_v7_pytest_dense.log; awk '!seen[$2 $3]++' gpurun_out/r02_v7_potrf_phases.log | head -30; tail -n 14 gpurun_out/r02_v7_pytest_sizes.log; cat gpurun_out/r02_v7_C3_9it.log; tail -n 8 gpurun_out/r02_v7_pytest_gpu.log
