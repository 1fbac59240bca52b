# full GPU suite on the build with the DMMA K tile / pipelined SYRK, finer phase breakdown, C3 bench line
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_i.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_i.log
bash tools/devrun2.sh libffvd_b200_dev.so > gpurun_out/r02_phase_fine.txt 2>&1; grep -v "^synth\|^actu\|^gas\|^drive" gpurun_out/r02_phase_fine.txt | head -20
python bench.py --steps 3 --warmup 3 --extras none > gpurun_out/r02_bench_c3_d.json 2> gpurun_out/r02_bench_c3_d.err; cut -c1-700 gpurun_out/r02_bench_c3_d.json
