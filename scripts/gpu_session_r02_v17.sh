mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_solver.py -q --maxfail=5 -m gpu 2>&1 | tail -12) > gpurun_out/r02_v17_pytest_solver.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -24) > gpurun_out/r02_v17_C3.log
(timeout 600 python scripts/op_profile.py C3 kkt_assemble 2>&1 | tail -14) > gpurun_out/r02_v17_op_profile_C3.log
tail -n 12 gpurun_out/r02_v17_pytest_solver.log; cat gpurun_out/r02_v17_C3.log; cat gpurun_out/r02_v17_op_profile_C3.log
