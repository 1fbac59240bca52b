#!/bin/bash
# A/B of two kernel-development builds on the GPU box: parity of each dev-minimal build, then the phase breakdowns
for L in "$@"; do
  export FFVD_B200_LIB=$PWD/ffvd_b200/lib/$L
  echo "=== $L"
  python tools/dev_check.py dev 2>&1 | tail -9 | cut -c1-220
  python tools/phase_timing.py 20000 256 8 16
  python tools/phase_timing.py 4000 512 16 8
  python tools/phase_timing.py 20000 100 4 16
done
