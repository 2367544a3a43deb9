#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v31}
timeout 900 python -m pytest tests/test_gpu_solver.py -x -q > gpurun_out/${TAG}_pytest_solver.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_solver.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "bigtop or e2500" > gpurun_out/${TAG}_pytest_big.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_big.log
timeout 600 python scripts/bench_dense.py 1000 2500 5000 10000 20000 > gpurun_out/${TAG}_dense.log 2>&1
tail -6 gpurun_out/${TAG}_dense.log
for R in 1 2; do
timeout 600 python bench.py > gpurun_out/${TAG}_bench$R.json 2> gpurun_out/${TAG}_bench$R.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench$R.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in list(d['kernel_ms_per_step'].items())[:8]})
print(d['roofline'])
"; tail -3 gpurun_out/${TAG}_bench$R.err
done
