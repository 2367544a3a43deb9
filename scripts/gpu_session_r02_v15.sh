mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py -q --maxfail=5 -m gpu -k "lanes or bigtop" 2>&1 | tail -8) > gpurun_out/r02_v15_pytest_lanes.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -24) > gpurun_out/r02_v15_C3_lanes8.log
(SMCP_B200_LANES=12 RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -24) > gpurun_out/r02_v15_C3_lanes12.log
tail -n 5 gpurun_out/r02_v15_pytest_lanes.log; for l in 8 12; do echo "== lanes $l"; cat gpurun_out/r02_v15_C3_lanes$l.log; done
