mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -30) > gpurun_out/r02_v10_pytest_multi.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r02_v10_bench_2gpu.err | tail -1) > gpurun_out/r02_v10_bench_2gpu.json
(RUNCFG_METHOD=feas RUNCFG_NOPROF=1 timeout 300 python scripts/run_config.py C4 full > gpurun_out/r02_v10_C4_feas.log 2>&1; echo "exit $?" >> gpurun_out/r02_v10_C4_feas.log)
tail -n 12 gpurun_out/r02_v10_pytest_multi.log; grep -E "NCCL INFO.*(nranks|Init COMPLETE)|smcp_b200: NCCL" gpurun_out/r02_v10_bench_2gpu.err | head -6; tail -n 3 gpurun_out/r02_v10_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v10_bench_2gpu.json'))
print("C3 N=2 e2e", d["e2e"]["value"], "value", d["value"], "launches", d["gpu_launches"])
print("schur_potrf", {k:v for k,v in d["schur_potrf"].items() if k not in ("note","flop_model")})
print("ops", {k: round(v["ms_per_step"],1) for k,v in d["chordal_ops_ms_per_step"].items()})
PY
tail -n 15 gpurun_out/r02_v10_C4_feas.log
