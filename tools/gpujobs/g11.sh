python -m pytest tests/test_gpu_parity.py -q -x -k "float32 or host_numpy or sghmc_kernel or adam or error_statuses or bounds" 2>&1 | tail -5
python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_plain_unc2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_final_unc python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_ncu_unc2.log 2>&1
python tools/run_one.py 4000 512 16 8 3 > gpurun_out/r02_plain_c5shape.log 2>&1 && \
FFVD_DL=1 ncu --set full --clock-control none -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_final_c5shape python tools/run_one.py 4000 512 16 8 3 > gpurun_out/r02_ncu_c5shape.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
