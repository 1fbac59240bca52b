"""Developer parity sweep: CUDA path (through the C ABI) vs the CPU oracle, per tensor.
Run on a GPU box:  python tools/dev_check.py [quick|large|dev]   (dev: uses lib/libffvd_b200_dev.so)"""
import copy
import sys
import os
if len(sys.argv) > 1 and sys.argv[1] == "dev":
    os.environ.setdefault("FFVD_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "ffvd_b200", "lib", "libffvd_b200_dev.so"))
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ffvd_b200
from oracle import fixtures, ffvd_oracle as O

dev = torch.device("cuda:0")
ctx = ffvd_b200.Context(0, torch.cuda.current_stream().cuda_stream)


def to_dev(a):
    return None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)


def run_cuda(prob, collapsed, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL):
    p = {k: to_dev(getattr(prob, k)) for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")}
    X = p["X"]
    S = 1 if X.dim() == 2 else X.shape[0]
    o = {"nll": torch.empty(S, dtype=torch.float64, device=dev), "terms": torch.empty(S, 6, dtype=torch.float64, device=dev)}
    for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
        if p[k] is not None:
            o["g_" + k] = torch.full_like(p[k], float("nan"))
    ctx.nll_grads(prob.kind, collapsed, p, o, flags=flags)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in o.items()}
    if S == 1 and X.dim() == 2:
        res["nll"] = res["nll"][0]; res["terms"] = res["terms"][0]
    return res


def compare(a, b, tag, tol=1e-9):
    worst = 0.0
    bad = []
    for k in a:
        if k not in b:
            continue
        x, y = np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)
        den = max(np.max(np.abs(x)), 1e-300)
        err = np.max(np.abs(x - y)) / den if np.all(np.isfinite(y)) else float("inf")
        worst = max(worst, err)
        if not (err <= tol):
            bad.append((k, err, den))
    print("%-58s worst rel err %.3e %s" % (tag, worst, "OK" if not bad else "FAIL " + str(bad)), flush=True)
    return not bad


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    devmin = len(sys.argv) > 1 and sys.argv[1] == "dev"      # -DFFVD_DEV_MINIMAL library: SE uncollapsed, Mp in {128,256,512}
    ok = True
    cases = []
    cases.append(("synthetic T=70 M=24 D=2 S=1", fixtures.synthetic_problem(T=70, M=24, D=2, S=1)))
    cases.append(("synthetic T=300 M=48 D=3 S=2", fixtures.synthetic_problem(T=300, M=48, D=3, S=2)))
    packed = fixtures.load_packed()
    byname = {p.name: p for p in packed["problems"]}
    for nm in ("actuator/0", "gas_furnace/0"):
        cases.append((nm, byname[nm]))
    p2 = copy.copy(byname["drive/0"]); p2.X = packed["extra"]["drive/0"]
    cases.append(("drive/0 S=4", p2))
    lin = copy.copy(byname["gas_furnace/0"]); lin.kind = 1; lin.logl = None; lin.logv = np.zeros(4)
    cases.append(("gas_furnace/0 LinearK", lin))
    if not quick:
        cases.append(("synthetic T=257 M=200 D=2 S=1 (Mp=256)", fixtures.synthetic_problem(T=257, M=200, D=2, S=1)))
        cases.append(("synthetic T=130 M=300 D=2 S=1 (Mp=384)", fixtures.synthetic_problem(T=130, M=300, D=2, S=1)))
        cases.append(("synthetic T=100 M=500 D=2 S=2 (Mp=512)", fixtures.synthetic_problem(T=100, M=500, D=2, S=2)))
        if len(sys.argv) > 1 and sys.argv[1] == "large":
            cases = []
            cases.append(("synthetic T=50 M=700 D=2 S=1 (Mp=768)", fixtures.synthetic_problem(T=50, M=700, D=2, S=1)))
            cases.append(("synthetic T=40 M=1000 D=2 S=1 (Mp=1024)", fixtures.synthetic_problem(T=40, M=1000, D=2, S=1)))
            cases.append(("synthetic T=30 M=1400 D=1 S=1 (Mp=1536)", fixtures.synthetic_problem(T=30, M=1400, D=1, S=1)))
            cases.append(("synthetic T=20 M=2048 D=1 S=1 (Mp=2048)", fixtures.synthetic_problem(T=20, M=2048, D=1, S=1)))
    if devmin:
        cases = [(t, p) for t, p in cases if p.kind == 0 and "Mp=384" not in t]
    for tag, prob in cases:
        for collapsed in ((False,) if devmin else (False, True)):
            t = time.time()
            ref = O.nll_and_grads(prob, collapsed=collapsed)
            try:
                got = run_cuda(prob, collapsed)
            except Exception as e:
                print("%-58s EXCEPTION %r" % (tag, str(e)[:300]), flush=True)
                return 1
            ok &= compare(ref, got, "%s collapsed=%d" % (tag, collapsed))
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
