# after hoisting the deterministic-mode lookups out of the RED sites + interleaved u-bar shuffles: tests, phases, plain timings
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_j.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_j.log
bash tools/devrun2.sh libffvd_b200_dev.so > gpurun_out/r02_phase_e.txt 2>&1; grep -v "^synth\|^actu\|^gas\|^drive" gpurun_out/r02_phase_e.txt | head -20
python tools/run_one.py 20000 256 8 16 3; python tools/run_one.py 4000 512 16 8 3; python tools/run_one.py 20000 100 4 16 3
