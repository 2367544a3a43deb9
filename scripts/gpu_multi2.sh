#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v32}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 scripts/multi_gpu_check.py 2000 1000 5 > gpurun_out/${TAG}_mgpu_check.log 2>&1
grep -E "rank|MULTI|Error|error" gpurun_out/${TAG}_mgpu_check.log | tail -5
timeout 600 $TR --master-port 29536 bench.py --gpus 2 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_n2.json'))
print('N=2 value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in list(d['kernel_ms_per_step'].items())[:8]})
"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_bench_n2.err | tail -3
timeout 300 $TR --master-port 29537 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref_n2.json 2> gpurun_out/${TAG}_bench_ref_n2.err
head -c 400 gpurun_out/${TAG}_bench_ref_n2.json
