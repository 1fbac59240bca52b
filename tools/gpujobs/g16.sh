python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_g.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_g.log
python tools/extra_bench.py > gpurun_out/r02_extra_bench_b.json 2> gpurun_out/r02_extra_bench_b.err; tail -3 gpurun_out/r02_extra_bench_b.err
python -c "
import json; r=json.load(open('gpurun_out/r02_extra_bench_b.json'))
for k,v in r.items():
    if k!='m_sweep_D8_S8': print(k, v)
for m,v in r['m_sweep_D8_S8'].items(): print(m, v)"
python tools/run_c1_graph.py > gpurun_out/r02_c1_graph_b.json 2>/dev/null; cat gpurun_out/r02_c1_graph_b.json
