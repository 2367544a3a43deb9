mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_v4_gpus.txt
(timeout 900 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -30) > gpurun_out/r02_v4_pytest_multi.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 scripts/bench_dense.py 10000 2>&1 | grep -E "^m=|rror" | tail -5) > gpurun_out/r02_v4_dense_2gpu.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r02_v4_bench_2gpu.err | tail -1) > gpurun_out/r02_v4_bench_2gpu.json
(timeout 400 python scripts/run_config.py C3 9 2>&1 | tail -34) > gpurun_out/r02_v4_C3_9it.log
tail -n 12 gpurun_out/r02_v4_pytest_multi.log; cat gpurun_out/r02_v4_dense_2gpu.log; grep -E "NCCL INFO.*(nranks|Init COMPLETE|comm )" gpurun_out/r02_v4_bench_2gpu.err | head -6; tail -n 3 gpurun_out/r02_v4_bench_2gpu.err; head -c 2500 gpurun_out/r02_v4_bench_2gpu.json; echo; head -14 gpurun_out/r02_v4_C3_9it.log
