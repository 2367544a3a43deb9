#!/bin/bash
# 2-GPU session: sharded assembly + block-cyclic Cholesky checks, dense microbench, bench at N=2
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v18}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 scripts/multi_gpu_check.py 2000 1000 5 > gpurun_out/${TAG}_mgpu_check.log 2>&1
grep -E "rank|MULTI|Error|error" gpurun_out/${TAG}_mgpu_check.log | tail -8
timeout 300 $TR --master-port 29534 scripts/multi_gpu_check.py 600 333 5 > gpurun_out/${TAG}_mgpu_check2.log 2>&1
grep -E "rank|MULTI|Error|error" gpurun_out/${TAG}_mgpu_check2.log | tail -8
timeout 600 $TR --master-port 29535 scripts/bench_dense.py 1000 5000 10000 20000 > gpurun_out/${TAG}_dense_n2.log 2>&1
grep -E "^m=|Error|error" gpurun_out/${TAG}_dense_n2.log | tail -8
timeout 600 python scripts/bench_dense.py 10000 > gpurun_out/${TAG}_dense_n1.log 2>&1
grep -E "^m=|Error|error" gpurun_out/${TAG}_dense_n1.log | tail -3
timeout 600 $TR --master-port 29536 bench.py --gpus 2 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
head -c 900 gpurun_out/${TAG}_bench_n2.json; tail -3 gpurun_out/${TAG}_bench_n2.err
