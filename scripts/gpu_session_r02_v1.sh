mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dense.py -x -q 2>&1 | tail -15) > gpurun_out/r02_v1_pytest_dense.log
(timeout 300 python scripts/bench_kernels.py 2>&1 | tail -40) > gpurun_out/r02_v1_bench_kernels.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02_v1_pytest_gpu.log
(timeout 400 python scripts/run_config.py C3 4 2>&1 | tail -30) > gpurun_out/r02_v1_C3.log
(timeout 300 python bench.py 2>gpurun_out/r02_v1_bench.err | tail -1) > gpurun_out/r02_v1_bench.json
tail -5 gpurun_out/r02_v1_pytest_dense.log gpurun_out/r02_v1_pytest_gpu.log; cat gpurun_out/r02_v1_bench_kernels.log; tail -22 gpurun_out/r02_v1_C3.log
