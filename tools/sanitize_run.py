"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
uncollapsed + collapsed nll/grads (SE and Linear, padded and ragged shapes, Din = 5 / 17), conditional with q_sqrt,
collapse_u_mean, SG-HMC and Adam updates.  usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ffvd_b200
from oracle import fixtures

dev = torch.device("cuda:0")
ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
KEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")


def run(prob, collapsed):
    p = {k: (None if getattr(prob, k) is None else torch.as_tensor(np.ascontiguousarray(getattr(prob, k)), dtype=torch.float64, device=dev)) for k in KEYS}
    S = 1 if p["X"].dim() == 2 else p["X"].shape[0]
    o = {"nll": torch.empty(S, dtype=torch.float64, device=dev), "terms": torch.empty(S, 6, dtype=torch.float64, device=dev)}
    for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
        if p[k] is not None:
            o["g_" + k] = torch.empty_like(p[k])
    ctx.nll_grads(prob.kind, collapsed, p, o)
    torch.cuda.synchronize()
    return p, o


cases = [dict(T=70, M=24, D=2, S=1), dict(T=97, M=130, D=3, S=2), dict(T=40, M=300, D=2, S=1), dict(T=33, M=500, D=16, S=1, n_ctrl=1),
         dict(T=20, M=700, D=2, S=1), dict(T=50, M=40, D=4, S=1, kind=1)]
for c in cases:
    prob = fixtures.synthetic_problem(**c)
    for collapsed in (False, True):
        p, o = run(prob, collapsed)
        print(c, "collapsed=%d nll=%.6f" % (collapsed, float(o["nll"][0])), flush=True)
# prediction operators
from ffvd_b200 import conditionals_multi_output as cmo
from ffvd_b200.kernels_multi_output import SquaredExponential
prob = fixtures.synthetic_problem(T=60, M=50, D=3, S=1)
kerns = [SquaredExponential(4, variance=np.exp(prob.logv[k]), lengthscales=np.exp(prob.logl[k]), ARD=True) for k in range(3)]
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
Xc = np.concatenate([prob.X[:60], prob.ctrl], axis=1)
U, L = cmo.collapse_u_mean_after_kernel_precalculation(None, t(Xc), t(prob.X), t(prob.Z), kerns, t(np.exp(prob.logQ)))
f = cmo.kernel_pre_cal(t(prob.Z), kerns)
mu, var = cmo.conditional_after_kernel_precalculation(f, t(Xc[:17]), t(prob.Z), kerns, U[0], white=True, q_sqrt=L)
mu2, var2 = cmo.conditional_after_kernel_precalculation(f, t(Xc[:5]), t(prob.Z), kerns, U[0], white=True, q_sqrt=L)
n = 1001
th = [torch.randn(n, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.ones(n, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.zeros(n, dtype=torch.float64, device=dev)]
ctx.sghmc_update(*th, 0.01, 0.05, 100.0, True)
ctx.sghmc_update(*th, 0.01, 0.05, 100.0, False)
ctx.adam_update(th[0], th[1], th[3], th[4], 1e-3, 0.9, 0.999, 1e-8, 1)
torch.cuda.synchronize()
print("done", float(mu.sum()), float(var2.sum()))
