mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_kernels.py -q --maxfail=10 -k "qr or trsm or operator_and_schur or golden" 2>&1 | tail -30) > gpurun_out/r02_v8_pytest_qr.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 6 2>&1 | tail -22) > gpurun_out/r02_v8_C3.log
(RUNCFG_NOPROF=1 RUNCFG_KKT=qr timeout 600 python scripts/run_config.py C4 full 2>&1 | tail -8) > gpurun_out/r02_v8_C4_esd_qr.log
(RUNCFG_NOPROF=1 timeout 600 python scripts/run_config.py C4 full 2>&1 | tail -8) > gpurun_out/r02_v8_C4_esd_chol.log
(timeout 900 python bench.py --workload band_n5000_m1000_bw5 --secondary none 2>gpurun_out/r02_v8_bench_c2.err | tail -1) > gpurun_out/r02_v8_bench_c2.json
tail -n 12 gpurun_out/r02_v8_pytest_qr.log; cat gpurun_out/r02_v8_C3.log; echo "--- C4 esd qr"; cat gpurun_out/r02_v8_C4_esd_qr.log; echo "--- C4 esd chol"; cat gpurun_out/r02_v8_C4_esd_chol.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v8_bench_c2.json'))
print("C2 e2e", d["e2e"]["value"], "value", d["value"], "launches", d["gpu_launches"]); print(d["kernel_ms_per_step"]); print(d["roofline"]); print(d["time_to_solve"])
PY
