python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_d.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_d.log
python tools/extra_bench.py > gpurun_out/r02_extra_bench_a.json 2> gpurun_out/r02_extra_bench_a.err; tail -3 gpurun_out/r02_extra_bench_a.err
python -c "
import json; r=json.load(open('gpurun_out/r02_extra_bench_a.json'))
for k,v in r.items():
    if k!='m_sweep_D8_S8': print(k, v)
for m,v in r['m_sweep_D8_S8'].items(): print(m, v)"
