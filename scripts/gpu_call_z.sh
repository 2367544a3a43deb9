#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v39}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29546 bench.py --gpus 4 > gpurun_out/${TAG}_bench_n4.json 2> gpurun_out/${TAG}_bench_n4.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench_n4.json') if l.startswith('{')][-1])
print('N=4 value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in list(d['kernel_ms_per_step'].items())[:8]})
"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_n4.err | tail -5; head -c 200 gpurun_out/${TAG}_bench_n4.json
