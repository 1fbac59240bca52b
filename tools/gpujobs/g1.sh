set -x
nvidia-smi -L
python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_plain_unc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r02_fused_base_unc python tools/run_one.py 20000 256 8 16 3 > gpurun_out/r02_ncu_unc.log 2>&1
python tools/run_one.py 20000 256 8 16 2 collapsed > gpurun_out/r02_plain_col.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fused_kernel|collapsed" -s 4 -c 4 -o gpurun_out/r02_fused_base_col python tools/run_one.py 20000 256 8 16 2 collapsed > gpurun_out/r02_ncu_col.log 2>&1
ls -la gpurun_out/*.ncu-rep
bash tools/devrun.sh > gpurun_out/r02_devrun_base.log 2>&1
FFVD_B200_LIB=$PWD/ffvd_b200/lib/libffvd_b200_dev.so python tools/phase_timing.py 20000 256 8 16 collapsed >> gpurun_out/r02_devrun_base.log 2>&1
tail -50 gpurun_out/r02_devrun_base.log
