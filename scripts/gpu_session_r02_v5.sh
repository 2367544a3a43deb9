mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -q --maxfail=8 -m gpu 2>&1 | tail -25) > gpurun_out/r02_v5_pytest_dense.log
(timeout 300 python scripts/bench_kernels.py gemm 2>&1 | tail -12) > gpurun_out/r02_v5_bench_kernels.log
(SMCP_B200_NO_TMA=1 timeout 300 python scripts/bench_kernels.py gemm 2>&1 | tail -12) > gpurun_out/r02_v5_bench_kernels_notma.log
(timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_dense.py 2>&1 | tail -30) > gpurun_out/r02_v5_pytest_gpu.log
(RUNCFG_NOPROF=1 timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -40) > gpurun_out/r02_v5_C3_9it.log
(timeout 900 python bench.py 2>gpurun_out/r02_v5_bench.err | tail -1) > gpurun_out/r02_v5_bench.json
tail -n 6 gpurun_out/r02_v5_pytest_dense.log; cat gpurun_out/r02_v5_bench_kernels.log; echo NOTMA; cat gpurun_out/r02_v5_bench_kernels_notma.log; tail -n 8 gpurun_out/r02_v5_pytest_gpu.log; cat gpurun_out/r02_v5_C3_9it.log; tail -n 3 gpurun_out/r02_v5_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v5_bench.json'))
print("C3 e2e", d["e2e"], "value", d["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"])
print("schur_potrf", {k:v for k,v in d["schur_potrf"].items() if k!="note"})
print("tts", d["time_to_solve"])
print("ops", d["chordal_ops_ms_per_step"])
s=d.get("secondary")
if s:
    print("C2 e2e", s["e2e"], "value", s["value"], "roofline", s["roofline"]); print(s["kernel_ms_per_step"]); print("tts", s["time_to_solve"])
PY
