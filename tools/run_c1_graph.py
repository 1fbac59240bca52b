"""BASELINE config 1 (one chain, actuator/0: T=512, M=100, D=4): one nll+gradient evaluation launch by launch, as a bound
call, and as a replayed CUDA graph; and a whole 21-evaluation sghmc_step launch by launch vs as ONE graph launch.
usage: python tools/run_c1_graph.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ffvd_b200
from ffvd_b200 import models
from ffvd_b200.datasets import load_packed_problems

dev = torch.device("cuda:0")
a = torch.randn(4096, 4096, device=dev)
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.0:          # leave the idle clocks
    (a @ a).sum().item()
ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
pr = {p["name"]: p for p in load_packed_problems(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))}["actuator/0"]
p = {k: torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for k, v in pr.items() if k != "name"}
o = {"nll": torch.empty(1, dtype=torch.float64, device=dev), "terms": torch.empty(1, 6, dtype=torch.float64, device=dev)}
for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
    o["g_" + k] = torch.empty_like(p[k])
FL = ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC


def best_ms(fn, reps=100):
    best = 1e9
    for _ in range(5):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / reps * 1e3)
    return best


res = {}
for collapsed in (False, True):
    tag = "collapsed" if collapsed else "uncollapsed"
    call = ctx.prepare_nll_grads(0, collapsed, p, o, flags=FL)
    res["eval_%s_bound_call_ms" % tag] = best_ms(call.run)
    g = ctx.capture(lambda: ctx.nll_grads(0, collapsed, p, o, flags=FL))
    res["eval_%s_graph_ms" % tag] = best_ms(g.launch)
    res["eval_%s_graph_kernels" % tag] = g.kernels
    g.close(); call.close()
# a whole sghmc_step (case 7: X and U sampled; case 2: kernel hyper-parameters and U sampled)
for case_val in (7, 2):
    args = dict(CC=pr["C"], DD=pr["d"], QQ_chol=np.exp(0.5 * pr["logQ"]), RR_chol=np.exp(pr["logR"]), lengthscales=np.exp(pr["logl"]),
                variance=np.exp(pr["logv"]), UU_ini=pr["U"], XX_0_ini=pr["X"][0], x_initialization=pr["X"][1:], ZZ=pr["Z"])
    for graph in (False, True):
        m = models.configure(models.RegressionModel("normal"), args, pr["ctrl"], case_val, iterations=0, window_size=64)
        model = m.fit(pr["Y"])
        model.enable_graph(graph)
        model.sghmc_step()
        res["sghmc_step_case%d_%s_ms" % (case_val, "graph" if graph else "launches")] = best_ms(model.sghmc_step, reps=10)
        if graph:
            res["sghmc_step_case%d_graph_kernels" % case_val] = model._graph.kernels
print(json.dumps(res, indent=1))
