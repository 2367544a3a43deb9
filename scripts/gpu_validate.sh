#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v37}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in list(d['kernel_ms_per_step'].items())[:8]})
print(d['roofline'])
print(d['time_to_solve'])
"; tail -3 gpurun_out/${TAG}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
