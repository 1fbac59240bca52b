python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/run_one.py 20000 256 8 16 3; python tools/run_one.py 4000 512 16 8 3; python tools/run_one.py 20000 100 4 16 3
