python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c3_n2_b.json 2> gpurun_out/r02_bench_c3_n2_b.err; echo rc=$?; tail -4 gpurun_out/r02_bench_c3_n2_b.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_c3_n2_b.json')); print(l['ms_per_step'], l['value'], l['roofline']['frac']); print(json.dumps(l['extra']['time_sharded_chain'], indent=0))"
