# collapsed mode (the CLI default) on the final kernel: ncu --set full of pass 1, collapsed_chol, collapsed_vec, pass 2 at the C3 shape; launch list
python tools/run_one.py 20000 256 8 16 3 collapsed > gpurun_out/r02_plain_col2.log 2>&1 && cat gpurun_out/r02_plain_col2.log && \
ncu --set full --clock-control none -k regex:"fused_kernel|collapsed" -s 8 -c 4 -o gpurun_out/r02_collapsed_final python tools/run_one.py 20000 256 8 16 3 collapsed > gpurun_out/r02_ncu_col2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/r02_launches_collapsed_final.csv python tools/run_one.py 20000 256 8 16 2 collapsed > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_collapsed_final.csv | tail -14
