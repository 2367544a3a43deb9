#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v23}
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_kernels.py -x -q -k "chain or bw5 or bw3 or bw7 or bw1 or bw2" > gpurun_out/${TAG}_pytest_chain.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_chain.log
timeout 600 python -m pytest tests/test_gpu_solver.py -x -q -k "driver or conelp" > gpurun_out/${TAG}_pytest_solver.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_solver.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
print(d['roofline'])
"; tail -3 gpurun_out/${TAG}_bench.err
