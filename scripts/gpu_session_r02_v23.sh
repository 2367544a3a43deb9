mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver.py -q --maxfail=8 -m gpu -x 2>&1 | tail -8) > gpurun_out/r02_v23_pytest_kernels.log
(timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -s --maxfail=4 -k "C3" 2>&1 | tail -6) > gpurun_out/r02_v23_pytest_sizes.log
L=gpurun_out/r02_v23_thresholds.log
: > $L
for cf in 6e6 6e5; do
  echo "== SMCP_B200_BIG_COMPL_FLOPS=$cf" >> $L
  (SMCP_B200_BIG_COMPL_FLOPS=$cf RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 6 2>&1 | grep -E "iteration [35]|op_|kkt_") >> $L
done
(SMCP_B200_BIG_COMPL_FLOPS=6e5 timeout 600 python scripts/op_profile.py C3 hessian hessian_inv kkt_assemble 2>&1 | tail -50) > gpurun_out/r02_v23_op_profile_C3.log
tail -n 4 gpurun_out/r02_v23_pytest_kernels.log; cat gpurun_out/r02_v23_pytest_sizes.log; cat $L; cat gpurun_out/r02_v23_op_profile_C3.log
