export FFVD_B200_LIB=$PWD/ffvd_b200/lib/libffvd_b200_dev.so
python tools/dev_check.py dev 2>&1 | tail -9 | cut -c1-200
python tools/phase_timing.py 20000 100 4 16
python tools/phase_timing.py 20000 64 4 16
python tools/phase_timing.py 20000 256 8 16 | head -3
