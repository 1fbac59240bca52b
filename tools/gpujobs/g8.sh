python -m pytest tests/test_gpu_parity.py -q -x -k "cuda_graph or collapsed_time or reuse_kzz or device_exp or logdensity_norm_full" 2>&1 | tail -5
python tools/run_c1_graph.py > gpurun_out/r02_c1_graph.json 2> gpurun_out/r02_c1_graph.err; tail -3 gpurun_out/r02_c1_graph.err; cat gpurun_out/r02_c1_graph.json
