# the same kernel source built four ways: does the phase-timing instrumentation (or a plain compiler fence at the marks) change the code?
for L in libffvd_b200.so libffvd_b200_devN.so libffvd_b200_devF.so libffvd_b200_dev.so; do
  echo "== $L"
  FFVD_B200_LIB=$PWD/ffvd_b200/lib/$L python tools/run_one.py 20000 256 8 16 3
  FFVD_B200_LIB=$PWD/ffvd_b200/lib/$L python tools/run_one.py 4000 512 16 8 3
  FFVD_B200_LIB=$PWD/ffvd_b200/lib/$L python tools/run_one.py 20000 100 4 16 3
done
