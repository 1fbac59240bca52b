set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_a.log 2>&1; tail -15 gpurun_out/r02_pytest_gpu_a.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c3_a.json 2> gpurun_out/r02_bench_c3_a.err; echo rc=$?; tail -5 gpurun_out/r02_bench_c3_a.err; cat gpurun_out/r02_bench_c3_a.json
