"""Secondary measurements (not the headline bench line): collapsed path, BASELINE config 4 (95 batched
chains), M sweep 64..2048 at roughly constant work, SG-HMC kernel HBM throughput.
usage (GPU box): python tools/extra_bench.py > gpurun_out/extra.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ffvd_b200
from bench import algorithmic_flops_per_unit, make_host_data
from oracle import fixtures

dev = torch.device("cuda:0")
ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
FL = ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC
GK = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR")


def dev_prob(h):
    return {k: (None if v is None else torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float64, device=dev)) for k, v in h.items()}


def outs_for(P):
    S = 1 if P["X"].dim() == 2 else P["X"].shape[0]
    o = {"nll": torch.empty(S, dtype=torch.float64, device=dev), "terms": torch.empty(S, 6, dtype=torch.float64, device=dev)}
    for k in GK:
        if P.get(k) is not None:
            o["g_" + k] = torch.empty_like(P[k])
    return o


def timeit(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
# ---- (1) uncollapsed vs collapsed at a C3-shaped size
T, M, D, S = 20000, 256, 8, 16
P = dev_prob(make_host_data(T, M, D, S, seed=1)); O = outs_for(P)
for collapsed in (False, True):
    ctx.fused_time(True)
    ms = timeit(lambda: ctx.nll_grads(0, collapsed, P, O, flags=FL))
    fms, fn = ctx.fused_time(True)
    res["c3like_T20k_M256_D8_S16_%s" % ("collapsed" if collapsed else "uncollapsed")] = dict(
        ms_per_eval=ms, units_per_s=S * T * D / ms * 1e3, fused_ms_per_eval=fms / fn * (2 if collapsed else 1),
        alg_tflops=(10 if collapsed else 6) * M * M * S * T * D / ms * 1e-9)
# ---- (2) BASELINE config 4: all 95 warm starts as one batched call
packed = fixtures.load_packed()["problems"]
hs = [dict(X=p.X, Z=p.Z, U=p.U, logv=p.logv, logl=p.logl, logQ=p.logQ, C=p.C, d=p.d, logR=p.logR, Y=p.Y, ctrl=p.ctrl) for p in packed]
Ps = [dev_prob(h) for h in hs]; Os = [outs_for(p) for p in Ps]
units = sum(p.Y.shape[0] * 4 for p in packed)
for collapsed in (False, True):
    ms_unbound = timeit(lambda: ctx.nll_grads_batched(0, collapsed, Ps, Os, flags=FL), reps=10, warm=3)
    call = ctx.prepare_nll_grads(0, collapsed, Ps, Os, flags=FL)          # tensors bound once (what a sampler loop does)
    ms = timeit(call.run, reps=20, warm=3)
    call.close()
    res["c4_95chains_%s" % ("collapsed" if collapsed else "uncollapsed")] = dict(ms_per_eval_all_chains=ms, chain_T_D_per_s=units / ms * 1e3,
                                                                                 evals_per_s=95 / ms * 1e3, ms_with_per_call_dlpack_export=ms_unbound)
for collapsed in (True, False):
    call = ctx.prepare_nll_grads(0, collapsed, Ps[0], Os[0], flags=FL)
    res["c1_single_chain_%s_ms" % ("collapsed" if collapsed else "uncollapsed")] = timeit(call.run, reps=50, warm=5)
    call.close()
# ---- (3) M sweep at ~constant algorithmic work (D=8, S=8)
sweep = {}
for M in (64, 128, 256, 512, 1024, 2048):
    T = max(256, int(2.0e12 / (6.0 * M * M * 8 * 8)))
    T = min(T, 60000)
    P = dev_prob(make_host_data(T, M, 8, 8, seed=2)); O = outs_for(P)
    ctx.fused_time(True)
    ms = timeit(lambda: ctx.nll_grads(0, False, P, O, flags=FL), reps=2, warm=1)
    fms, fn = ctx.fused_time(True)
    units = 8 * T * 8
    sweep[str(M)] = dict(T=T, ms_per_eval=ms, fused_ms=fms / fn, units_per_s=units / ms * 1e3,
                         alg_tflops_fused=algorithmic_flops_per_unit(M, 9) * units / (fms / fn) * 1e-9)
    del P, O
res["m_sweep_D8_S8"] = sweep
# ---- (4) SG-HMC kernel HBM throughput (96 B/element burn-in, 56 B sample)
n = 1 << 28
ten = [torch.randn(n, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.ones(n, dtype=torch.float64, device=dev) for _ in range(3)] + \
      [torch.zeros(n, dtype=torch.float64, device=dev)]
for burn, nbytes in ((True, 96), (False, 56)):
    ms = timeit(lambda: ctx.sghmc_update(*ten, 0.01, 0.05, 1e5, burn), reps=5, warm=2)
    res["sghmc_%s" % ("burn_in" if burn else "sample")] = dict(elements=n, ms=ms, gb_per_s=n * nbytes / ms * 1e-6)
print(json.dumps(res, indent=1))
