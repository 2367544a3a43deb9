#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v24}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_solver.py -x -q > gpurun_out/${TAG}_pytest_solver.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_solver.log
timeout 600 python scripts/bench_dense.py 1000 2500 10000 > gpurun_out/${TAG}_dense.log 2>&1
tail -4 gpurun_out/${TAG}_dense.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
print(d['roofline'])
"; tail -3 gpurun_out/${TAG}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/profile_iter.py 5000 1000 5 3 > gpurun_out/${TAG}_ncu_list.log 2>&1
python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
gzip -f gpurun_out/${TAG}_launches.csv
head -45 gpurun_out/${TAG}_launches_summary.txt
