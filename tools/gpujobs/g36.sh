# final sanity of the round: GPU suite, smoke, default bench line, ncu --set full of the final kernel (C3 shape + configs[4] shape)
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_m.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_m.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_c3_i.json 2> gpurun_out/r02_bench_c3_i.err; python -c "
import json; r=json.load(open('gpurun_out/r02_bench_c3_i.json'))
print(r['ms_per_step'], r['value'], r['e2e']['value'], r['roofline']['frac'], r['roofline']['traffic'], r['parity']['cuda_vs_oracle']['worst'], r['gpu_launches'])"
python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_plain_unc7.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_final4_unc python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_ncu_unc7.log 2>&1
python tools/run_one.py 4000 512 16 8 3 > gpurun_out/r02_plain_c5shape3.log 2>&1 && \
ncu --set full --clock-control none -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_final4_c5shape python tools/run_one.py 4000 512 16 8 3 > gpurun_out/r02_ncu_c5shape3.log 2>&1
cat gpurun_out/r02_plain_unc7.log gpurun_out/r02_plain_c5shape3.log
