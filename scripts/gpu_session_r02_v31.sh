mkdir -p gpurun_out
L=gpurun_out/r02_v31_potrf_la.log
: > $L
for la in 0 24 48 148; do
  echo "== SMCP_B200_POTRF_LA=$la" >> $L
  (SMCP_B200_POTRF_LA=$la timeout 300 python scripts/bench_dense.py 4000 10000 2>&1 | grep -E "m= *(4000|10000|20000)") >> $L
done
T=gpurun_out/r02_v31_thresholds.log
: > $T
for cf in 6e5 2e5 6e4; do
  echo "== SMCP_B200_BIG_COMPL_FLOPS=$cf" >> $T
  (SMCP_B200_BIG_COMPL_FLOPS=$cf RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [356]|op_completion|op_hessian |op_hessian_inv|op_cholesky|prep_inv") >> $T
done
cat $L $T
