mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q --maxfail=10 -m gpu 2>&1 | tail -8) > gpurun_out/r02_v37_pytest_gpu.log
(timeout 900 python bench.py 2>gpurun_out/r02_v37_bench.err | tail -1) > gpurun_out/r02_v37_bench.json
(timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>gpurun_out/r02_v37_bench_ref.err | tail -1) > gpurun_out/r02_v37_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_v37_smoke.log 2>&1
(timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_v37_launches.csv python scripts/op_profile.py C3 completion hessian hessian_inv cholesky kkt_assemble kkt_factor > gpurun_out/r02_v37_ncu_list.log 2>&1)
python scripts/summarize_launches.py gpurun_out/r02_v37_launches.csv > gpurun_out/r02_v37_launches_summary.txt 2>&1
gzip -f gpurun_out/r02_v37_launches.csv
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:potrf_tile -s 6 -c 1 -f -o gpurun_out/r02_v37_ncu_potrf_tile python scripts/op_profile.py C3 completion > gpurun_out/r02_v37_ncu_potrf_tile.log 2>&1)
python scripts/ncu_summary.py gpurun_out/r02_v37_ncu_potrf_tile.ncu-rep > gpurun_out/r02_v37_ncu_potrf_tile.txt 2>&1
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_up -s 20 -c 1 -f -o gpurun_out/r02_v37_ncu_thin_up python scripts/op_profile.py C3 hessian > gpurun_out/r02_v37_ncu_thin_up.log 2>&1)
python scripts/ncu_summary.py gpurun_out/r02_v37_ncu_thin_up.ncu-rep > gpurun_out/r02_v37_ncu_thin_up.txt 2>&1
tail -n 4 gpurun_out/r02_v37_pytest_gpu.log; tail -2 gpurun_out/r02_v37_bench.err; tail -2 gpurun_out/r02_v37_smoke.log; head -c 600 gpurun_out/r02_v37_bench_ref.json; echo
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v37_bench.json'))
print("C3 e2e", d["e2e"], "value", d["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"])
print("schur_potrf", {k:v for k,v in d["schur_potrf"].items() if k not in ("note","flop_model")})
print("tts", d["time_to_solve"])
print("ops", {k: round(v["ms_per_step"],1) for k,v in d["chordal_ops_ms_per_step"].items()})
print("kern", {k: round(v,1) for k,v in list(d["kernel_ms_per_step"].items())[:16]})
print("fam", {k: (v["bound"], round(v["frac"],3) if v["frac"] else None) for k,v in d["family_rooflines"].items()})
s=d.get("secondary")
if s:
    print("C2 e2e", s["e2e"], "value", s["value"], "roofline", s["roofline"]["kernel"], s["roofline"]["frac"]); print("tts", s["time_to_solve"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
head -30 gpurun_out/r02_v37_launches_summary.txt
grep -E "kernel:|gpu__time_duration|dram__bytes|dram_throughput|warps_active|dmma_cycles|stalled_(long|barrier|wait|short)" gpurun_out/r02_v37_ncu_potrf_tile.txt gpurun_out/r02_v37_ncu_thin_up.txt | head -40
