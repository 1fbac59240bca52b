python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_plain_unc6.log 2>&1 && cat gpurun_out/r02_plain_unc6.log && \
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_final3_unc python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_ncu_unc6.log 2>&1
ls -la gpurun_out/r02_fused_final3_unc.ncu-rep
