# ncu source-level capture of the fused kernel with the DMMA K tile (C3 shape, 16 samples)
python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_plain_unc3.log 2>&1 && cat gpurun_out/r02_plain_unc3.log && \
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_kmma_unc python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_ncu_unc3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
