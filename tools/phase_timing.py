"""Per-phase cycle breakdown of the fused uncollapsed kernel (diagnostic build, -DFFVD_PHASE_TIMING).
usage: python tools/phase_timing.py T M D S [collapsed]
Build first:  ./build.sh -DFFVD_PHASE_TIMING -o ffvd_b200/lib/libffvd_b200_prof.so   (build.sh's last -o wins)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FFVD_B200_LIB", os.path.join(ROOT, "ffvd_b200", "lib", "libffvd_b200_prof.so"))
sys.path.insert(0, ROOT)
import torch
import ffvd_b200
from bench import make_host_data

NAMES = ["flush+decode+stage x", "K tile (r2, exp)", "A = K L^-T (+row partials)", "store A, row stats, emission",
         "ubar + SYRK S+=A^T A (warp 0)", "Kbar = Abar L^-1", "W = Kbar o K, store", "contract: rows of xbar (+tail)",
         "contract: W^T X~ -> Zbar (warp 0)", "contract: W Z~ (warp 0)", "contract: barrier wait",
         "(P0) wait for the previous d / x tile", "(P3) barrier + store A", "(P6) K from scratch, W in registers", "(P6) barrier wait", "(P3) row statistics on warp 0 (phase 3 is then the wait for the other warps)"]
T, M, D, S = map(int, sys.argv[1:5])
collapsed = len(sys.argv) > 5 and sys.argv[5] == "collapsed"
dev = torch.device("cuda:0")
ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
h = make_host_data(T, M, D, S, seed=1)
P = {k: torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for k, v in h.items()}
out = {"nll": torch.empty(S, dtype=torch.float64, device=dev), "terms": torch.empty(S, 6, dtype=torch.float64, device=dev)}
for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
    out["g_" + k] = torch.empty_like(P[k])
fl = ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC
for _ in range(2):
    ctx.nll_grads(0, collapsed, P, out, flags=fl)
ctx.phase_clocks(True)
ctx.fused_time(True)
ctx.nll_grads(0, collapsed, P, out, flags=fl)
ms, n = ctx.fused_time(True)
clk = ctx.phase_clocks(True)
tot = sum(clk)
print("T=%d M=%d D=%d S=%d collapsed=%d fused %.3f ms over %d launch(es)" % (T, M, D, S, collapsed, ms, n))
for i, nm in enumerate(NAMES):
    print("  phase %d %-34s %6.2f %%   %8.0f clk/item/CTA-avg" % (i, nm, 100.0 * clk[i] / tot, clk[i] / max(1, (D * S * ((T + 63) // 64)))))
