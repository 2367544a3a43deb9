mkdir -p gpurun_out
L=gpurun_out/r02_v20_thresholds.log
: > $L
for cf in 6e6 2e6 6e5 2e5; do
  echo "== SMCP_B200_BIG_COMPL_FLOPS=$cf" >> $L
  (SMCP_B200_BIG_COMPL_FLOPS=$cf RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 6 2>&1 | grep -E "iteration [35]|op_completion|op_hessian |op_hessian_inv|op_cholesky|prep_inv") >> $L
done
for bf in 6e5 2e5; do
  echo "== SMCP_B200_BIG_FLOPS=$bf SMCP_B200_BIG_COMPL_FLOPS=6e5" >> $L
  (SMCP_B200_BIG_FLOPS=$bf SMCP_B200_BIG_COMPL_FLOPS=6e5 RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 6 2>&1 | grep -E "iteration [35]|op_completion|op_hessian |op_hessian_inv|op_cholesky|prep_inv") >> $L
done
cat $L
