# 8-GPU run of the final round-2 build: 8-rank correctness check, then the bench with all extras (C3 weak + strong, configs[4] at FULL T = 1M with S = 256 over 8 GPUs, 95 chains dealt over 8 ranks, time-sharded chain)
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/check_multigpu.py > gpurun_out/r02_multigpu_check_n$N.txt 2>&1; tail -4 gpurun_out/r02_multigpu_check_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_n${N}_b.json 2> gpurun_out/r02_bench_c3_n${N}_b.err; echo rc=$?; tail -4 gpurun_out/r02_bench_c3_n${N}_b.err
python -c "
import json,sys; l=json.load(open('gpurun_out/r02_bench_c3_n${N}_b.json')); print(l['ms_per_step'], l['value'], l['e2e']['value'], l['roofline']['frac'], l['parity'].get('multi_gpu_allreduce',{}).get('worst'))
for k in ('c3_strong','c5'):
    e=l['extra'].get(k)
    if e: print(k, e['ms_per_step'], e['value'], e['executed_frac_of_peak'], e['fused_ms_per_rank'])
print(json.dumps(l['extra'].get('c4_95chains'))[:600]); print(json.dumps(l['extra'].get('time_sharded_chain'))[:800])"
