mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q --maxfail=10 -m gpu 2>&1 | tail -12) > gpurun_out/r02_v24_pytest_gpu.log
(timeout 900 python bench.py 2>gpurun_out/r02_v24_bench.err | tail -1) > gpurun_out/r02_v24_bench.json
tail -n 8 gpurun_out/r02_v24_pytest_gpu.log; tail -3 gpurun_out/r02_v24_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v24_bench.json'))
print("C3 e2e", d["e2e"], "value", d["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"])
print("schur_potrf", {k:v for k,v in d["schur_potrf"].items() if k not in ("note","flop_model")})
print("tts", d["time_to_solve"])
print("ops", {k: round(v["ms_per_step"],1) for k,v in d["chordal_ops_ms_per_step"].items()})
print("kern", {k: round(v,1) for k,v in list(d["kernel_ms_per_step"].items())[:16]})
s=d.get("secondary")
if s:
    print("C2 e2e", s["e2e"], "value", s["value"], "roofline", s["roofline"]); print({k: round(v,3) for k,v in list(s["kernel_ms_per_step"].items())[:12]}); print("tts", s["time_to_solve"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
