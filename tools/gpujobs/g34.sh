python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_l.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_l.log
python tools/run_one.py 20000 256 8 16 3; python tools/run_one.py 4000 512 16 8 3; python tools/run_one.py 20000 100 4 16 3
python bench.py > gpurun_out/r02_bench_c3_g.json 2> gpurun_out/r02_bench_c3_g.err; cut -c1-300 gpurun_out/r02_bench_c3_g.json
python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c3_g.json'))
print(r['ms_per_step'], r['value'], r['e2e']['value'], r['roofline']['frac'], r['roofline']['traffic'], r['parity']['cuda_vs_oracle']['worst'])"
