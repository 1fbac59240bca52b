"""Time one uncollapsed evaluation (small T, so the K(Z,Z) preparation dominates) per factorisation path:
default choice vs FFVD_BLOCKED_CHOL=1 (multi-kernel blocked path) vs FFVD_BLOCKED_CHOL=0 (single-CTA path).
usage: python tools/prep_paths.py [M ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ffvd_b200
from oracle import fixtures

dev = torch.device("cuda:0")
ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
Ms = [int(a) for a in sys.argv[1:]] or [100, 119, 120, 128, 144, 160, 161, 200, 256]
KEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")
a = torch.randn(4096, 4096, device=dev)
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.0:
    (a @ a).sum().item()
for M in Ms:
    prob = fixtures.synthetic_problem(T=64, M=M, D=4, S=1, seed=M)
    p = {k: (None if getattr(prob, k) is None else torch.as_tensor(np.ascontiguousarray(getattr(prob, k)), dtype=torch.float64, device=dev)) for k in KEYS}
    o = {"nll": torch.empty(1, dtype=torch.float64, device=dev), "terms": torch.empty(1, 6, dtype=torch.float64, device=dev)}
    for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
        o["g_" + k] = torch.empty_like(p[k])
    call = ctx.prepare_nll_grads(prob.kind, False, p, o, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC)
    res = []
    for mode in (None, "1", "0"):
        if mode is None:
            os.environ.pop("FFVD_BLOCKED_CHOL", None)
        else:
            os.environ["FFVD_BLOCKED_CHOL"] = mode
        try:
            for _ in range(5):
                call.run()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    call.run()
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 20)
            res.append("%.3f ms" % best)
        except Exception as ex:
            res.append("failed (%s)" % str(ex)[:40])
    os.environ.pop("FFVD_BLOCKED_CHOL", None)
    print("M %4d (T=64, D=4): default %s   blocked %s   single-CTA %s" % (M, res[0], res[1], res[2]), flush=True)
