# dual half-width CTAs with a start offset for the second CTA of every SM (0 / 60 / 120 percent of Mp^2 clocks)
for L in libffvd_b200_devF.so libffvd_b200_devN.so libffvd_b200_devB.so; do
  echo "== $L"
  for D in 0 1; do
  FFVD_DUAL=$D FFVD_B200_LIB=$PWD/ffvd_b200/lib/$L python tools/run_one.py 20000 256 8 16 3
  FFVD_DUAL=$D FFVD_B200_LIB=$PWD/ffvd_b200/lib/$L python tools/run_one.py 20000 100 4 16 3
  done
done
