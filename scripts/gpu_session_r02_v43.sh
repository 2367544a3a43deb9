mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -q --maxfail=10 -m gpu 2>&1 | tail -6) > gpurun_out/r02_v43_pytest_gpu.log
(timeout 600 python bench.py 2>gpurun_out/r02_v43_bench.err | tail -1) > gpurun_out/r02_v43_bench.json
tail -n 4 gpurun_out/r02_v43_pytest_gpu.log; tail -2 gpurun_out/r02_v43_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v43_bench.json'))
print("C3 e2e", d["e2e"], "value", d["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
print("tts", d["time_to_solve"])
print("fam", {k: (v["bound"], round(v["frac"],3) if v["frac"] else None) for k,v in d["family_rooflines"].items()})
print("schur_potrf", {k:v for k,v in d["schur_potrf"].items() if k not in ("note","flop_model","regions_ms_per_step")})
s=d.get("secondary")
print("C2 e2e", s["e2e"], "value", s["value"]); print("tts", s["time_to_solve"])
PY
