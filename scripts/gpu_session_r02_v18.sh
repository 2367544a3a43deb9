mkdir -p gpurun_out
(timeout 900 python bench.py 2>gpurun_out/r02_v18_bench.err | tail -1) > gpurun_out/r02_v18_bench.json
(RUNCFG_NOPROF=1 RUNCFG_PER_ITER=1 timeout 400 python scripts/run_config.py C3 12 2>&1 | tail -40) > gpurun_out/r02_v18_C3.log
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:scm_kstream -c 1 -f -o gpurun_out/r02_v18_ncu_scm_kstream python scripts/op_profile.py C3 kkt_assemble > gpurun_out/r02_v18_ncu_scm_kstream.log 2>&1)
python scripts/ncu_summary.py gpurun_out/r02_v18_ncu_scm_kstream.ncu-rep > gpurun_out/r02_v18_ncu_scm_kstream.txt 2>&1
(timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 120000 --csv --log-file gpurun_out/r02_v18_launches.csv python bench.py --steps 2 --warmup 1 --no-solve --secondary none > gpurun_out/r02_v18_ncu_list.log 2>&1)
python scripts/summarize_launches.py gpurun_out/r02_v18_launches.csv > gpurun_out/r02_v18_launches_summary.txt 2>&1
gzip -f gpurun_out/r02_v18_launches.csv
tail -3 gpurun_out/r02_v18_bench.err; cat gpurun_out/r02_v18_C3.log; head -40 gpurun_out/r02_v18_launches_summary.txt
grep -E "kernel:|gpu__time_duration|dram__bytes|dram_throughput|sm__warps_active" gpurun_out/r02_v18_ncu_scm_kstream.txt | head
