# smoke(), reference arm, and the DRAM traffic of ONE fused launch of the C3 bench workload (dram__bytes only: two replays)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | cut -c1-300
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fused_kernel -s 6 -c 1 --csv --log-file gpurun_out/r02_c3_traffic.csv python bench.py --steps 2 --warmup 3 --extras none > gpurun_out/r02_c3_traffic.log 2>&1
grep -v "^==" gpurun_out/r02_c3_traffic.csv | cut -c1-400 | tail -5
