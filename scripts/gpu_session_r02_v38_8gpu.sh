mkdir -p gpurun_out
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_v38_bench_n$N.json 2> gpurun_out/r02_v38_bench_n$N.err
  echo "N=$N rc=$?"
  grep -E "NCCL INFO ncclCommInitRank|nranks|Traceback|Error" gpurun_out/r02_v38_bench_n$N.err | head -12 > gpurun_out/r02_v38_bench_n$N.nccl.txt
  tail -c 2000 gpurun_out/r02_v38_bench_n$N.err > gpurun_out/r02_v38_bench_n$N.err.tail; rm -f gpurun_out/r02_v38_bench_n$N.err
done
(timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -5) > gpurun_out/r02_v38_pytest_multi.log
python - <<'PY'
import json
for N in (8,4,2):
    try:
        d=json.load(open('gpurun_out/r02_v38_bench_n%d.json'%N))
        print(N, "e2e", round(d["e2e"]["value"],4), "value", round(d["value"],4), {k: round(v["ms_per_step"],1) for k,v in d["schur_potrf"]["regions_ms_per_step"].items()}, {k: round(v["ms_per_step"],1) for k,v in d["chordal_ops_ms_per_step"].items()})
    except Exception as e:
        print(N, "failed", e)
PY
cat gpurun_out/r02_v38_pytest_multi.log; tail -3 gpurun_out/r02_v38_bench_n8.err.tail
