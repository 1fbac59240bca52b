"""Run n evaluations of the fused nll+grads path on one synthetic problem (for ncu captures).
usage: python tools/run_one.py T M D S [n] [collapsed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ffvd_b200
from bench import make_host_data
T, M, D, S = map(int, sys.argv[1:5])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 3
collapsed = len(sys.argv) > 6 and sys.argv[6] == "collapsed"
dev = torch.device("cuda:0")
ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
h = make_host_data(T, M, D, S, seed=1)
P = {k: torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for k, v in h.items()}
out = {"nll": torch.empty(S, dtype=torch.float64, device=dev), "terms": torch.empty(S, 6, dtype=torch.float64, device=dev)}
for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
    out["g_" + k] = torch.empty_like(P[k])
fl = ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC
for _ in range(n):
    ctx.nll_grads(0, collapsed, P, out, flags=fl)
ms, cnt = ctx.fused_time(True)
print("fused %.3f ms/launch over %d launches; nll sum %.12g" % (ms / max(cnt, 1), cnt, float(out["nll"].sum())))
