# full GPU suite on the constexpr-Mp build + ncu source capture (C3 shape) + C3 bench line
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_i.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_i.log
python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_plain_unc4.log 2>&1 && cat gpurun_out/r02_plain_unc4.log && \
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_kmma2_unc python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_ncu_unc4.log 2>&1
python bench.py --steps 3 --warmup 3 --extras none > gpurun_out/r02_bench_c3_d.json 2> gpurun_out/r02_bench_c3_d.err; cut -c1-600 gpurun_out/r02_bench_c3_d.json
