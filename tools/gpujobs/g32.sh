# after ltu_kernel over row chunks + triangular k ranges in the batched GEMM: full suite + large-M timings
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import subprocess,sys
for T,M in ((1241,2048),(4967,1024),(19868,512),(20000,256)):
    subprocess.run([sys.executable,"tools/run_one.py",str(T),str(M),"8","8","3"])
    subprocess.run([sys.executable,"tools/run_one.py",str(T),str(M),"8","8","3","collapsed"])
PY
python tools/extra_bench.py > gpurun_out/r02_extra_bench_d.json 2> gpurun_out/r02_extra_bench_d.err
python -c "
import json; r=json.load(open('gpurun_out/r02_extra_bench_d.json'))
for k,v in r.items():
    if k!='m_sweep_D8_S8': print(k, v)
for m,v in r['m_sweep_D8_S8'].items(): print(m, v)"
