python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_f.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_f.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | cut -c1-400
