python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_e.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu_e.log
python bench.py --steps 3 --warmup 3 --extras none > gpurun_out/r02_bench_c3_c.json 2> gpurun_out/r02_bench_c3_c.err; echo rc=$?; tail -3 gpurun_out/r02_bench_c3_c.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_c3_c.json')); print(l['ms_per_step'], l['value'], l['roofline']['frac'], l['parity']['cuda_vs_oracle']['worst'])"
