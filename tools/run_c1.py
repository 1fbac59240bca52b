"""BASELINE config 1: one chain (actuator/0 warm start), repeated nll+gradient evaluations through a bound call.
usage: python tools/run_c1.py [collapsed|x] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ffvd_b200
from oracle import fixtures
collapsed = len(sys.argv) > 1 and sys.argv[1] == "collapsed"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")


def warm_clocks(seconds=1.0):
    """sub-millisecond timings are meaningless on a GPU that has not left its idle clocks"""
    import time as _t
    a = torch.randn(4096, 4096, device=dev)
    t0 = _t.perf_counter()
    while _t.perf_counter() - t0 < seconds:
        (a @ a).sum().item()


ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
prob = {p.name: p for p in fixtures.load_packed()["problems"]}["actuator/0"]
KEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")
p = {k: torch.as_tensor(np.ascontiguousarray(getattr(prob, k)), dtype=torch.float64, device=dev) for k in KEYS}
o = {"nll": torch.empty(1, dtype=torch.float64, device=dev), "terms": torch.empty(1, 6, dtype=torch.float64, device=dev)}
for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
    o["g_" + k] = torch.empty_like(p[k])
call = ctx.prepare_nll_grads(0, collapsed, p, o, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC)
warm_clocks()
best = 1e9
for rep in range(5):
    for _ in range(5):
        call.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        call.run()
    torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / reps * 1e3)
print("actuator/0 (T=512, M=100, D=4), collapsed=%d: %.3f ms per evaluation (best of 5 x %d)" % (collapsed, best, reps))
