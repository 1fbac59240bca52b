# 95 chains (configs[3]) in one batched call: device time, with REUSE_KZZ, and the per-launch durations (warm caches)
python tools/run_c4.py x 20; python tools/run_c4.py x 20 reuse; python tools/run_c4.py collapsed 20
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/r02_launches_c4_final.csv python tools/run_c4.py x 3 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_c4_final.csv | tail -16
