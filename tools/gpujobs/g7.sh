python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_c.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py > gpurun_out/r02_multigpu_check_n2.txt 2>&1; tail -3 gpurun_out/r02_multigpu_check_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c3_n2_a.json 2> gpurun_out/r02_bench_c3_n2_a.err; echo rc=$?; tail -5 gpurun_out/r02_bench_c3_n2_a.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_c3_n2_a.json')); print(l['ms_per_step'], l['value'], l['roofline']['frac'], json.dumps(l['parity'])[:700]); print(json.dumps(l['extra'],indent=0)[:3500])"
