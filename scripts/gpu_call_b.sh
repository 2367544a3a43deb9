#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v15}
timeout 600 python -m pytest tests/test_gpu_solver.py -x -q -k "potrf or schur" > gpurun_out/${TAG}_pytest_dense.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_dense.log
timeout 600 python scripts/bench_dense.py 1000 2500 5000 10000 20000 > gpurun_out/${TAG}_dense.log 2>&1
tail -8 gpurun_out/${TAG}_dense.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/'+__import__('sys').argv[1]+'_bench.json')) if False else None
PY
python -c "
import json,sys
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
print('e2e',d['e2e'])
"
