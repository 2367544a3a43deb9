mkdir -p gpurun_out
(timeout 200 python bench.py --steps 2 --warmup 3 --no-solve --secondary none 2>gpurun_out/r02_v45_bench.err | tail -1) > gpurun_out/r02_v45_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v45_bench.json'))
print("C3 e2e", d["e2e"]["value"], "value", d["value"])
c=d["cpu_baseline"]; print("cpu", c.get("value"), c.get("cores"), c.get("error")); print((c.get("sample") or "")[:420])
PY
tail -2 gpurun_out/r02_v45_bench.err
