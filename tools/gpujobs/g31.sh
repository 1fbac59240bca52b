# large-M preparation: per-launch durations at M = 2048 and M = 1024 (D = 8)
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/r02_launches_m2048.csv python tools/run_one.py 1241 2048 8 8 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_m2048.csv | tail -16
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/r02_launches_m1024.csv python tools/run_one.py 4967 1024 8 8 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_m1024.csv | tail -16
