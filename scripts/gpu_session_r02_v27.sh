mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_solver.py -q --maxfail=8 -m gpu -k "dense_potrf or feas or schur" 2>&1 | tail -8) > gpurun_out/r02_v27_pytest_solver.log
(timeout 300 python scripts/bench_dense.py 2>&1 | tail -12) > gpurun_out/r02_v27_dense.log
(RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [35678]|op_|kkt_|status") > gpurun_out/r02_v27_C3.log
(SMCP_B200_POTRS_NO_WAVE=1 RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 6 2>&1 | grep -E "iteration [35]|kkt_solve") > gpurun_out/r02_v27_C3_nowave.log
tail -n 5 gpurun_out/r02_v27_pytest_solver.log; cat gpurun_out/r02_v27_dense.log gpurun_out/r02_v27_C3.log gpurun_out/r02_v27_C3_nowave.log
