# round 2, session 3: DMMA K tile + deterministic flag -- dev parity + phase breakdown, then the full GPU suite
bash tools/devrun.sh > gpurun_out/r02_phase_ktile_mma.txt 2>&1; cat gpurun_out/r02_phase_ktile_mma.txt | head -50
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_h.log 2>&1; tail -15 gpurun_out/r02_pytest_gpu_h.log
