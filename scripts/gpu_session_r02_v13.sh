mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -q --maxfail=8 -m gpu 2>&1 | tail -15) > gpurun_out/r02_v13_pytest_dense.log
(SMCP_B200_PT_DEBUG=1 timeout 300 python scripts/bench_kernels.py potrf 2>&1 | grep -E "^potrf|m=(1000|1186|1131|2000|2560) ") > gpurun_out/r02_v13_potrf_phases.log
(timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_dense.py --deselect tests/test_gpu_baseline_sizes.py 2>&1 | tail -12) > gpurun_out/r02_v13_pytest_gpu.log
(timeout 600 python scripts/op_profile.py C3 completion hessian hessian_inv 2>&1 | tail -40) > gpurun_out/r02_v13_op_profile_C3.log
(timeout 900 python bench.py 2>gpurun_out/r02_v13_bench.err | tail -1) > gpurun_out/r02_v13_bench.json
tail -n 5 gpurun_out/r02_v13_pytest_dense.log; awk '!seen[$2 $3]++' gpurun_out/r02_v13_potrf_phases.log | head -24; tail -n 6 gpurun_out/r02_v13_pytest_gpu.log; cat gpurun_out/r02_v13_op_profile_C3.log; tail -3 gpurun_out/r02_v13_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v13_bench.json'))
print("C3 e2e", d["e2e"], "value", d["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"])
print("schur_potrf", {k:v for k,v in d["schur_potrf"].items() if k not in ("note","flop_model")})
print("tts", d["time_to_solve"])
print("ops", {k: round(v["ms_per_step"],1) for k,v in d["chordal_ops_ms_per_step"].items()})
print("kern", {k: round(v,1) for k,v in list(d["kernel_ms_per_step"].items())[:12]})
s=d.get("secondary")
if s:
    print("C2 e2e", s["e2e"], "value", s["value"], "roofline", s["roofline"]); print({k: round(v,3) for k,v in list(s["kernel_ms_per_step"].items())[:12]}); print("tts", s["time_to_solve"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
