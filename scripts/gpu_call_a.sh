#!/bin/bash
# GPU-box session A: parity tests, bench, ncu launch list, the other BASELINE configs for a few iterations.
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v14}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
head -c 800 gpurun_out/${TAG}_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/profile_iter.py 5000 1000 5 3 > gpurun_out/${TAG}_ncu_list.log 2>&1
python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
gzip -f gpurun_out/${TAG}_launches.csv
timeout 300 python scripts/run_config.py C4 4 > gpurun_out/${TAG}_C4.log 2>&1; tail -20 gpurun_out/${TAG}_C4.log
timeout 500 python scripts/run_config.py C3 3 2000 10000 > gpurun_out/${TAG}_C3.log 2>&1; tail -20 gpurun_out/${TAG}_C3.log
timeout 300 python scripts/run_config.py C5 3 5000 > gpurun_out/${TAG}_C5_n5000.log 2>&1; tail -20 gpurun_out/${TAG}_C5_n5000.log
ls -la gpurun_out
