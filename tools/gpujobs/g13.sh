python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_n4_a.json 2> gpurun_out/r02_bench_c3_n4_a.err; echo rc=$?; tail -4 gpurun_out/r02_bench_c3_n4_a.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_c3_n4_a.json')); print(l['ms_per_step'], l['value'], l['e2e']['value'], l['roofline']['frac'], l['parity']['multi_gpu_allreduce']['worst'])
for k in ('c3_strong','c5'):
    e=l['extra'][k]; print(k, e['ms_per_step'], e['value'], e['executed_frac_of_peak'], e['fused_ms_per_rank'])
print(l['extra']['c4_95chains']['uncollapsed'], l['extra']['c4_95chains']['collapsed'])"
