#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v16}
timeout 900 python -m pytest tests/test_gpu_solver.py -x -q > gpurun_out/${TAG}_pytest_solver.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_solver.log
timeout 600 python scripts/bench_dense.py 1000 2500 5000 10000 20000 > gpurun_out/${TAG}_dense.log 2>&1
tail -8 gpurun_out/${TAG}_dense.log
for LA in 1 0; do
  SMCP_B200_LOOKAHEAD=$LA timeout 600 python bench.py > gpurun_out/${TAG}_bench_la$LA.json 2> gpurun_out/${TAG}_bench_la$LA.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_la$LA.json'))
print('lookahead=$LA value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
"
done
timeout 900 python scripts/run_config.py C3 3 2000 10000 > gpurun_out/${TAG}_C3.log 2>&1; tail -22 gpurun_out/${TAG}_C3.log
