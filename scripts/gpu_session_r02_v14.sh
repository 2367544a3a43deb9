mkdir -p gpurun_out
# concurrent lanes over the top set: correctness (forced tiny fronts, 3 lanes), then C3 regions with 1 / 4 lanes
(timeout 600 python -m pytest tests/test_gpu_kernels.py -q --maxfail=5 -m gpu -k "lanes or bigtop3" 2>&1 | tail -8) > gpurun_out/r02_v14_pytest_lanes.log
(SMCP_B200_LANES=1 RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -24) > gpurun_out/r02_v14_C3_lanes1.log
(SMCP_B200_LANES=4 RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -24) > gpurun_out/r02_v14_C3_lanes4.log
(SMCP_B200_LANES=8 RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -24) > gpurun_out/r02_v14_C3_lanes8.log
tail -n 5 gpurun_out/r02_v14_pytest_lanes.log; for l in 1 4 8; do echo "== lanes $l"; cat gpurun_out/r02_v14_C3_lanes$l.log; done
