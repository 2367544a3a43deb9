mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_solver.py -q --maxfail=8 -m gpu -k "potrf or gemm or schur" 2>&1 | tail -4) > gpurun_out/r02_v35_pytest_dense.log
(timeout 300 python scripts/bench_dense.py 4000 10000 20000 2>&1 | grep -E "m= ") > gpurun_out/r02_v35_dense.log
(SMCP_B200_POTRF_NO_PT=1 timeout 300 python scripts/bench_dense.py 10000 20000 2>&1 | grep -E "m= ") > gpurun_out/r02_v35_dense_nopt.log
(RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [3568]|kkt_|status") > gpurun_out/r02_v35_C3.log
tail -n 3 gpurun_out/r02_v35_pytest_dense.log; cat gpurun_out/r02_v35_dense.log; echo nopt; cat gpurun_out/r02_v35_dense_nopt.log gpurun_out/r02_v35_C3.log
