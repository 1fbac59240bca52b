#!/bin/bash
# kernel-development loop on the GPU box: parity of the dev-minimal build, then the phase breakdown
export FFVD_B200_LIB=$PWD/ffvd_b200/lib/libffvd_b200_dev.so
python tools/dev_check.py dev 2>&1 | tail -12 | cut -c1-220
python tools/phase_timing.py 20000 256 8 16
python tools/phase_timing.py 4000 512 16 8
python tools/phase_timing.py 20000 100 4 16
