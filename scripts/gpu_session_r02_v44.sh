mkdir -p gpurun_out
(timeout 150 ncu --set full --clock-control none --import-source on -k regex:potrs_wave -s 1 -c 1 -f -o gpurun_out/r02_v44_ncu_potrs_wave python scripts/bench_dense.py 1000 > gpurun_out/r02_v44_ncu_potrs_wave.log 2>&1)
python scripts/ncu_summary.py gpurun_out/r02_v44_ncu_potrs_wave.ncu-rep > gpurun_out/r02_v44_ncu_potrs_wave.txt 2>&1
grep -E "kernel:|gpu__time_duration|dram__bytes|lts__t_bytes|warps_active|stalled_(long|barrier|wait|short|membar)|grid_size" gpurun_out/r02_v44_ncu_potrs_wave.txt | head -20
