export FFVD_B200_LIB=$PWD/ffvd_b200/lib/libffvd_b200_dev.so
python -m pytest tests/test_gpu_parity.py -q -x -k "device_exp or logdensity_norm_full or particle_gibbs" 2>&1 | tail -5
python tools/dev_check.py dev 2>&1 | tail -8 | cut -c1-200
python tools/phase_timing.py 20000 256 8 16
FFVD_DL=1 python tools/phase_timing.py 4000 512 16 8
python tools/phase_timing.py 20000 100 4 16
