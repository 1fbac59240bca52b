"""Measured CUDA-vs-oracle error over every bundled fixture: 95 Factnonlin_ini warm starts x {uncollapsed, collapsed},
max-norm relative per tensor (the tolerance convention of the tests).  Run on a GPU box:
    python tools/parity_report.py > profiles/r01_parity.txt"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ffvd_b200
from oracle import fixtures, ffvd_oracle as O

dev = torch.device("cuda:0")
ctx = ffvd_b200.Context(0, torch.cuda.current_stream().cuda_stream)
packed = fixtures.load_packed()
KEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")
worst = {}
per_ds = {}
for prob in packed["problems"]:
    p = {k: (None if getattr(prob, k) is None else torch.as_tensor(np.ascontiguousarray(getattr(prob, k)), dtype=torch.float64, device=dev)) for k in KEYS}
    for collapsed in (False, True):
        o = {"nll": torch.empty(1, dtype=torch.float64, device=dev), "terms": torch.empty(1, 6, dtype=torch.float64, device=dev)}
        for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
            o["g_" + k] = torch.empty_like(p[k])
        ctx.nll_grads(prob.kind, collapsed, p, o)
        torch.cuda.synchronize()
        ref = O.nll_and_grads(prob, collapsed=collapsed)
        for k, v in o.items():
            r = np.asarray(ref[k], dtype=np.float64).reshape(-1)
            g = v.cpu().numpy().reshape(-1)
            den = max(np.max(np.abs(r)), 1e-300)
            err = float(np.max(np.abs(r - g)) / den)
            key = (k, collapsed)
            worst[key] = max(worst.get(key, 0.0), err)
            ds = prob.name.split("/")[0]
            per_ds[(ds, collapsed)] = max(per_ds.get((ds, collapsed), 0.0), err)
print("# CUDA path (libffvd_b200.so through the C ABI) vs CPU oracle, B200, %d fixtures x 2 modes" % len(packed["problems"]))
print("# error = max|cuda - oracle| / max|oracle| per tensor; tolerance in tests: 1e-9")
print("%-10s %14s %14s" % ("tensor", "uncollapsed", "collapsed"))
for k in ("nll", "terms", "g_X", "g_Z", "g_U", "g_logv", "g_logl", "g_logQ", "g_C", "g_d", "g_logR"):
    print("%-10s %14.3e %14.3e" % (k, worst[(k, False)], worst[(k, True)]))
print()
print("%-12s %14s %14s" % ("dataset", "uncollapsed", "collapsed"))
for ds in sorted({d for d, _ in per_ds}):
    print("%-12s %14.3e %14.3e" % (ds, per_ds[(ds, False)], per_ds[(ds, True)]))
print()
print("overall worst: %.3e" % max(worst.values()))
