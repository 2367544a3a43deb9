mkdir -p gpurun_out
L=gpurun_out/r02_v29_lanes.log
: > $L
for nl in 8 12 16; do
  echo "== SMCP_B200_LANES=$nl" >> $L
  (SMCP_B200_LANES=$nl RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C3 9 2>&1 | grep -E "iteration [3568]|op_completion|op_hessian_inv|op_hessian_prep_inv|op_cholesky") >> $L
done
cat $L
