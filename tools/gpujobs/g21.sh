# dual half-width CTA experiment repeated on the kernel with the DMMA K tile + pipelined SYRK flush
for D in 0 1; do
  echo "FFVD_DUAL=$D"
  FFVD_DUAL=$D python tools/run_one.py 20000 256 8 16 3
  FFVD_DUAL=$D python tools/run_one.py 20000 100 4 16 3
done
