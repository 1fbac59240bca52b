# round-2 final kernel: full GPU suite, plain timings, ncu --set full captures (C3 shape with source, configs[4] shape), bench line, ncu launch list of the bench command
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_k.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_k.log
python tools/run_one.py 20000 256 8 16 3; python tools/run_one.py 4000 512 16 8 3; python tools/run_one.py 20000 100 4 16 3
python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_plain_unc5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_final2_unc python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_ncu_unc5.log 2>&1
python tools/run_one.py 4000 512 16 8 3 > gpurun_out/r02_plain_c5shape2.log 2>&1 && \
ncu --set full --clock-control none -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_final2_c5shape python tools/run_one.py 4000 512 16 8 3 > gpurun_out/r02_ncu_c5shape2.log 2>&1
python bench.py > gpurun_out/r02_bench_c3_f.json 2> gpurun_out/r02_bench_c3_f.err; cut -c1-300 gpurun_out/r02_bench_c3_f.json
python bench.py --steps 2 --warmup 3 --extras none > gpurun_out/r02_bench_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3_final.csv python bench.py --steps 2 --warmup 3 --extras none > gpurun_out/r02_ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_c3_final.csv | tail -4
