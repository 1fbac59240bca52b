"""BASELINE config 4: all 95 Factnonlin_ini warm starts as ONE batched nll+gradient call (for ncu launch lists).
usage: python tools/run_c4.py [collapsed|x] [reps] [reuse]   (reuse: FFVD_FLAG_REUSE_KZZ, Z / hyper-parameters fixed)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ffvd_b200
from oracle import fixtures
collapsed = len(sys.argv) > 1 and sys.argv[1] == "collapsed"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")


def warm_clocks(seconds=1.0):
    """sub-millisecond timings are meaningless on a GPU that has not left its idle clocks"""
    import time as _t
    a = torch.randn(4096, 4096, device=dev)
    t0 = _t.perf_counter()
    while _t.perf_counter() - t0 < seconds:
        (a @ a).sum().item()


ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
packed = fixtures.load_packed()
KEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")
Ps, Os = [], []
for prob in packed["problems"]:
    p = {k: (None if getattr(prob, k) is None else torch.as_tensor(np.ascontiguousarray(getattr(prob, k)), dtype=torch.float64, device=dev)) for k in KEYS}
    o = {"nll": torch.empty(1, dtype=torch.float64, device=dev), "terms": torch.empty(1, 6, dtype=torch.float64, device=dev)}
    for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
        o["g_" + k] = torch.empty_like(p[k])
    Ps.append(p); Os.append(o)
FL = ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC
if len(sys.argv) > 3 and sys.argv[3] == "reuse":
    FL |= ffvd_b200.FLAG_REUSE_KZZ
call = ctx.prepare_nll_grads(0, collapsed, Ps, Os, flags=FL)
warm_clocks()
for _ in range(5):
    call.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    call.run()
e1.record(); torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(reps):
    call.run()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / reps * 1e3
print("flags=%d " % FL, end="")
print("95 chains, collapsed=%d: %.3f ms per batched evaluation (device), %.3f ms wall" % (collapsed, e0.elapsed_time(e1) / reps, wall))
