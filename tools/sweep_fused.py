"""Time the fused uncollapsed kernel for the tile configurations named on the command line.
usage: python tools/sweep_fused.py T M D S  rb:nw [rb:nw ...]"""
import os, subprocess, sys, json
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import ffvd_b200
    from bench import make_host_data, algorithmic_flops_per_unit
    T, M, D, S = map(int, sys.argv[2:6])
    dev = torch.device("cuda:0")
    ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
    h = make_host_data(T, M, D, S, seed=1)
    P = {k: torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for k, v in h.items()}
    out = {"nll": torch.empty(S, dtype=torch.float64, device=dev), "terms": torch.empty(S, 6, dtype=torch.float64, device=dev)}
    for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
        out["g_" + k] = torch.empty_like(P[k])
    fl = ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC
    for _ in range(2):
        ctx.nll_grads(0, False, P, out, flags=fl)
    ctx.fused_time(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ctx.nll_grads(0, False, P, out, flags=fl)
    e1.record(); torch.cuda.synchronize()
    ms, n = ctx.fused_time(True)
    units = S * T * D
    print(json.dumps(dict(cfg=os.environ.get("FFVD_RB", "") + ":" + os.environ.get("FFVD_NW", ""), fused_ms=ms / n, total_ms=e0.elapsed_time(e1) / 3,
                          alg_tflops=algorithmic_flops_per_unit(M, D + 1) * units / (ms / n) * 1e-9, nll=float(out["nll"].sum()),
                          gZ=float(out["g_Z"].abs().sum()), gX=float(out["g_X"].abs().sum()))))
else:
    T, M, D, S = sys.argv[1:5]
    for cfg in sys.argv[5:]:
        rb, nw = cfg.split(":")
        env = dict(os.environ, FFVD_RB=rb, FFVD_NW=nw)
        r = subprocess.run([sys.executable, __file__, "--child", T, M, D, S], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
