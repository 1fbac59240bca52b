# round-2 final-ish build: C3 bench line (default flags = what the driver runs), extra bench (M sweep, collapsed, 95 chains, C1), C1 graph timings
python bench.py > gpurun_out/r02_bench_c3_e.json 2> gpurun_out/r02_bench_c3_e.err; cut -c1-400 gpurun_out/r02_bench_c3_e.json
python tools/extra_bench.py > gpurun_out/r02_extra_bench_c.json 2> gpurun_out/r02_extra_bench_c.err; tail -3 gpurun_out/r02_extra_bench_c.err
python -c "
import json; r=json.load(open('gpurun_out/r02_extra_bench_c.json'))
for k,v in r.items():
    if k!='m_sweep_D8_S8': print(k, v)
for m,v in r['m_sweep_D8_S8'].items(): print(m, v)"
python tools/run_c1_graph.py > gpurun_out/r02_c1_graph_c.json 2>/dev/null; cat gpurun_out/r02_c1_graph_c.json
