#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list, ncu full captures of the top kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-r01_v3}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -c "import json,os; print(open('MEASURED_PEAKS.json').read() if os.path.exists('MEASURED_PEAKS.json') else 'no MEASURED_PEAKS.json')" > gpurun_out/${TAG}_peaks.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | head -c 3000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/profile_iter.py 5000 1000 5 3 > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -30 gpurun_out/${TAG}_ncu_list.log
for K in chain_down_kernel chain_up_kernel chain_scan_kernel gemm_dmma_kernel potrs_kernel potrf_panel_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 4 \
      -f -o gpurun_out/${TAG}_$K python scripts/profile_iter.py 5000 1000 5 1 > gpurun_out/${TAG}_ncu_$K.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${TAG}_$K.ncu-rep > gpurun_out/${TAG}_ncu_$K.txt 2>&1
  ncu -i gpurun_out/${TAG}_$K.ncu-rep --page source --csv > gpurun_out/${TAG}_src_$K.csv 2>/dev/null
  gzip -f gpurun_out/${TAG}_src_$K.csv
  rm -f gpurun_out/${TAG}_$K.ncu-rep gpurun_out/${TAG}_ncu_$K.log
done
ls -la gpurun_out
