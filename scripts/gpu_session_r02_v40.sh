mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_dense.py tests/test_golden.py tests/test_gpu_chain.py -q --maxfail=8 -m gpu 2>&1 | tail -5) > gpurun_out/r02_v40_pytest.log
(timeout 900 python bench.py 2>gpurun_out/r02_v40_bench.err | tail -1) > gpurun_out/r02_v40_bench.json
tail -n 4 gpurun_out/r02_v40_pytest.log; tail -2 gpurun_out/r02_v40_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v40_bench.json'))
print("C3 e2e", d["e2e"], "value", d["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"])
print("tts", d["time_to_solve"])
print("fam", {k: (v["bound"], round(v["frac"],3) if v["frac"] else None) for k,v in d["family_rooflines"].items()})
s=d.get("secondary")
print("C2 e2e", s["e2e"], "value", s["value"], "roofline", s["roofline"]["kernel"], s["roofline"]["frac"]); print("tts", s["time_to_solve"])
print({k: round(v,3) for k,v in list(s["kernel_ms_per_step"].items())[:10]})
PY
