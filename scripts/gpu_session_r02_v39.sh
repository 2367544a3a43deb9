mkdir -p gpurun_out
L=gpurun_out/r02_v39_potrs_wave_min.log
: > $L
for wm in 4097 512; do
  echo "== SMCP_B200_POTRS_WAVE_MIN=$wm" >> $L
  (SMCP_B200_POTRS_WAVE_MIN=$wm timeout 200 python scripts/bench_dense.py 1000 2500 4000 2>&1 | grep -E "m= ") >> $L
  (SMCP_B200_POTRS_WAVE_MIN=$wm RUNCFG_NOPROF=1 timeout 200 python scripts/run_config.py C2 12 2>&1 | grep -E "iteration (5|7|9)|kkt_solve") >> $L
done
cat $L
