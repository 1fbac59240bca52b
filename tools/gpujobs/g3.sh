export FFVD_B200_LIB=$PWD/ffvd_b200/lib/libffvd_b200_dev.so
for dl in 1 0; do
echo "=== FFVD_DL=$dl"
FFVD_DL=$dl python tools/dev_check.py dev 2>&1 | tail -12 | cut -c1-200
FFVD_DL=$dl python tools/phase_timing.py 20000 256 8 16
FFVD_DL=$dl python tools/phase_timing.py 4000 512 16 8
FFVD_DL=$dl python tools/phase_timing.py 20000 100 4 16
done
