#!/usr/bin/env python
"""bench.py -- FFVD GPSSM transitions+grads/sec per SG-HMC step on B200 (BASELINE.json metric).

A "step" is ONE evaluation of the GPSSM nll and all its gradients (uncollapsed q(u), SE kernel,
float64) on synthetic data of the named shape, the (N>1) all-reduce of the shared-parameter
gradients, and ONE adaptive SG-HMC burn-in update of the sampled set {X, U} (case-7 style)
with injected noise.  Units = S*T*D transitions per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c3small]

N>1 is launched by the driver under torchrun (one rank per GPU, samples sharded S/rank = 64 each,
weak scaling).  --impl reference times the CPU restatement of the reference graph (oracle/) on a
bounded sample of the same workload on the box's host cores (TensorFlow itself is not installable).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

_REAL_STDOUT = sys.stdout

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the largest single-GPU configuration (configs[4], T=1M/M=512/D=16/S=256,
    # needs >= 2 GPUs for its 197 GB of X / x-bar / SG-HMC state -- SURVEY 7.2)
    "c3": dict(T=100_000, M=256, D=8, S=64, name="synthetic GPSSM T=100k M=256 D=8 S=64 SE float64 (BASELINE configs[2])"),
    "c3small": dict(T=10_000, M=256, D=8, S=8, name="synthetic GPSSM T=10k M=256 D=8 S=8 SE float64 (smoke-sized)"),
    # BASELINE.json configs[4] (T=1M, M=512, D=16, S=256 over 8 GPUs = 32 trajectories per GPU): the per-GPU shard, and the
    # same shard with T cut to 100k so that a step takes seconds instead of ~19 s (not the headline line: --workload only)
    "c5": dict(T=1_000_000, M=512, D=16, S=32, name="synthetic GPSSM T=1M M=512 D=16 S=256/8 per GPU SE float64 (BASELINE configs[4] shard)"),
    "c5short": dict(T=100_000, M=512, D=16, S=32, name="synthetic GPSSM T=100k M=512 D=16 S=32 per GPU SE float64 (configs[4] shape, T cut 10x)"),
}
CPU_SAMPLE_T = 20_000          # bounded CPU sample: T=20k, S=1 of the same M, D


def algorithmic_flops_per_unit(M, Din):
    return 6 * M * M + 9 * M * Din + 8 * M          # SURVEY 8(d), dense convention


def executed_flops_per_unit(M, Din):
    """Flops the fused kernel actually issues per (s,t,d) unit: both triangular contractions skip
    the zero half of L^{-1} (16-column groups), the SYRK forms only the 8x8 blocks on or below the diagonal of its 32x32
    tiles, plus the two thin products of the W back-propagation and the FP64 scalar work of the K tile."""
    Mp = next(s for s in (128, 256, 384, 512, 768, 1024, 1536, 2048) if M <= s)
    G, nt, nbx = Mp // 16, Mp // 32, (Din + 1 + 7) // 8
    tri = 128 * G * (G + 1)                                   # MACs per row of one triangular contraction
    syrk = 64 * (16 * (nt * (nt - 1) // 2) + 10 * nt)         # MACs per row of S += a a^T
    thin = 2 * Mp * 8 * nbx                                   # W^T [X,1] and W [Z,1]
    scalar = (Din + 25) * Mp                                  # r^2, exp, statistics
    return 2 * (2 * tri + syrk + thin) + scalar


def make_host_data(T, M, D, S, seed):
    """SURVEY 8(d) synthetic recipe (AR(1) trajectories, N(0,1.5^2) inducing inputs, ...), on the host."""
    from scipy.signal import lfilter
    rng = np.random.default_rng(seed)
    Din = D + 1
    ctrl = rng.standard_normal((T, 1))
    eps = rng.standard_normal((S, T + 1, D))
    eps[:, 1:] *= np.sqrt(1 - 0.95 ** 2)
    X = lfilter([1.0], [1.0, -0.95], eps, axis=1)             # x_{t+1} = 0.95 x_t + e_t, x_0 ~ N(0,1)
    shared = np.random.default_rng(20230209)                  # shared parameters identical on every rank
    Z = shared.standard_normal((M, Din)) * 1.5
    logl = np.log(shared.uniform(1.0, 4.0, (D, Din)))
    logv = np.log(shared.uniform(0.05, 0.8, D))
    logQ = 2.0 * np.log(shared.uniform(0.2, 0.8, D))
    logR = np.log(np.full((1, 1), 0.4))
    C = shared.standard_normal((D, 1)) * 0.3
    d = np.zeros(1)
    U = shared.standard_normal((M, D))
    Y = X[0, 1:] @ C + d + 0.4 * rng.standard_normal((T, 1))
    return dict(X=np.ascontiguousarray(X), Z=Z, U=U, logv=logv, logl=logl, logQ=logQ, C=C, d=d, logR=logR, Y=Y, ctrl=ctrl)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def time_cpu_oracle(cfg, steps, warmup):
    """CPU restatement of the reference graph (oracle/, torch CPU float64 + autograd) on a bounded sample."""
    import torch
    from oracle import ffvd_oracle as O
    from oracle.ffvd_oracle import Problem
    # all the host threads available (torchrun exports OMP_NUM_THREADS=1 to its workers: undo that for the CPU arm)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    T = min(CPU_SAMPLE_T, cfg["T"])
    h = make_host_data(T, cfg["M"], cfg["D"], 1, seed=1)
    prob = Problem(X=h["X"][0], Z=h["Z"], U=h["U"], logv=h["logv"], logl=h["logl"], logQ=h["logQ"], C=h["C"], d=h["d"],
                   logR=h["logR"], Y=h["Y"], ctrl=h["ctrl"])
    rng = np.random.default_rng(0)
    st = {n: [np.ones_like(getattr(prob, n)), np.ones_like(getattr(prob, n)), np.ones_like(getattr(prob, n)), np.zeros_like(getattr(prob, n))]
          for n in ("X", "U")}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        res = O.nll_and_grads(prob, collapsed=False)
        for n in ("X", "U"):
            xi, g, g2, p = st[n]
            th, xi, g, g2, p = O.sghmc_update(getattr(prob, n), res["g_" + n], rng.standard_normal(getattr(prob, n).shape), xi, g, g2, p,
                                              epsilon=0.01, mdecay=0.05, X_N=T + 1, burn_in=True)
            st[n] = [xi, g, g2, p]
            setattr(prob, n, th)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    units = T * cfg["D"]
    return dict(value=units / float(np.mean(times)), unit="transitions+grads/s", cores=torch.get_num_threads(), kind="port",
                sample="oracle (torch CPU float64 restatement of the TF graph + autograd) on T=%d, S=1, M=%d, D=%d of the same synthetic "
                       "workload, %d timed steps, %.2f s/step; host has %d logical cpus" % (T, cfg["M"], cfg["D"], len(times), float(np.mean(times)), os.cpu_count()),
                ms_per_step=1e3 * float(np.mean(times)))


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    cb = time_cpu_oracle(cfg, steps, warmup)
    line = {"impl": "reference", "metric": "GP transitions+grads/sec (S*T*D) per SG-HMC step", "value": cb["value"],
            "unit": "transitions+grads/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "sample": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "transitions+grads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was re-routed to stderr."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                                   # native libraries (NCCL) print banners to fd 1
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, cfg, rank, world)

    import torch
    import torch.distributed as dist
    import ffvd_b200
    from ffvd_b200 import distributed as fd
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    T, M, D, S = cfg["T"], cfg["M"], cfg["D"], cfg["S"]
    Din = D + 1
    ctx = ffvd_b200.Context(local, torch.cuda.current_stream(local).cuda_stream)

    # ---- data: host (pinned) copies for the e2e leg, device-resident copies for `value`
    h = make_host_data(T, M, D, S, seed=1000 + rank)
    P = {k: torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for k, v in h.items()}
    X_host = torch.as_tensor(h["X"]).pin_memory()
    torch.manual_seed(4321 + rank)                  # fixed noise: nll_mean is comparable from run to run
    noise_host = torch.randn(X_host.shape, dtype=torch.float64).pin_memory()
    noise_X = noise_host.to(dev)                    # pre-filled device noise: generation is outside the timed region
    noise_U = torch.randn(P["U"].shape, dtype=torch.float64, device=dev)
    out = {"nll": torch.empty(S, dtype=torch.float64, device=dev), "terms": torch.empty(S, 6, dtype=torch.float64, device=dev)}
    for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
        out["g_" + k] = torch.empty_like(P[k])
    state = {n: dict(xi=torch.ones_like(P[n]), g=torch.ones_like(P[n]), g2=torch.ones_like(P[n]), p=torch.zeros_like(P[n])) for n in ("X", "U")}
    X0, U0 = P["X"].clone(), P["U"].clone()
    flags = ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC

    def step(Pd=None, nzX=None):
        Pd = P if Pd is None else Pd
        nzX = noise_X if nzX is None else nzX
        ctx.nll_grads(ffvd_b200.KERNEL_SE, False, Pd, out, flags=flags)
        if world > 1:
            fd.allreduce_shared(out)                # one packed NCCL all-reduce of Z/U/hyper gradients
        for n, nz in (("X", nzX), ("U", noise_U)):
            st = state[n]
            ctx.sghmc_update(Pd[n], out["g_" + n], nz, st["xi"], st["g"], st["g2"], st["p"], 0.01, 0.05, float(T + 1), True)

    def reset():
        P["X"].copy_(X0); P["U"].copy_(U0)
        for n in ("X", "U"):
            state[n]["xi"].fill_(1.0); state[n]["g"].fill_(1.0); state[n]["g2"].fill_(1.0); state[n]["p"].zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- measured FP64 peak (cuBLAS DGEMM) for the roofline denominator, outside the timed region
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev); b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    best = 1e30
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); (a @ b); e1.record(); torch.cuda.synchronize()
        if i:
            best = min(best, e0.elapsed_time(e1))
    fp64_peak_tflops = 2 * 8192 ** 3 / best * 1e-9
    del a, b

    # ---- value: inputs resident in HBM
    for _ in range(W):
        step()
    reset()
    barrier()
    ctx.fused_time(reset=True)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count - launches0
    fused_ms, fused_n = ctx.fused_time(reset=True)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / K
    # per-rank device time of the fused kernel (diagnostic: shows which GPU of the box set the max)
    per_rank_fused = [fused_ms / max(fused_n, 1)]
    if world > 1:
        tl = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(tl, torch.tensor([per_rank_fused[0]], dtype=torch.float64, device=dev))
        per_rank_fused = [float(t.item()) for t in tl]
    units = S * T * D * world
    value = units / (ms_per_step * 1e-3)
    nll_check = float(out["nll"].mean().item())

    # ---- e2e: host buffers, H2D of the step's inputs and D2H of its result inside the timed region
    reset()
    nll_host = torch.empty(S, dtype=torch.float64).pin_memory()

    # Every step's inputs (X and its noise, 819 MB at C3) come from pinned host memory and every step's result goes back,
    # all inside the timed region.  The copies of step k+1 travel on a copy stream into a second pair of device buffers
    # while step k computes (the first step's copy is exposed); the host still waits for every step's nll.
    main_stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream(device=dev)
    Pb = [P, dict(P)]
    Pb[1]["X"] = torch.empty_like(P["X"])
    Nb = [noise_X, torch.empty_like(noise_X)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[i])         # the step that last read buffer pair i has finished
            Pb[i]["X"].copy_(X_host, non_blocking=True)
            Nb[i].copy_(noise_host, non_blocking=True)
            ready[i].record(copy_stream)

    def run_e2e(n):
        prefetch(0)
        for k in range(n):
            i = k & 1
            main_stream.wait_event(ready[i])
            if k + 1 < n:
                prefetch(1 - i)
            step(Pb[i], Nb[i])
            free[i].record(main_stream)
            nll_host.copy_(out["nll"], non_blocking=True)
            main_stream.synchronize()

    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    run_e2e(K)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    tmax = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    e2e_value = units / (float(tmax.item()) / K * 1e-3)
    h2d = X_host.numel() * 8 + noise_host.numel() * 8
    d2h = S * 8

    # ---- the second kernel of the step against its own roofline: the SG-HMC burn-in update of X is an HBM
    #      stream of 96 B / element (7 reads + 5 writes)
    reset()
    st = state["X"]
    for _ in range(2):
        ctx.sghmc_update(P["X"], out["g_X"], noise_X, st["xi"], st["g"], st["g2"], st["p"], 0.01, 0.05, float(T + 1), True)
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(5):
        ctx.sghmc_update(P["X"], out["g_X"], noise_X, st["xi"], st["g"], st["g2"], st["p"], 0.01, 0.05, float(T + 1), True)
    u1.record(); torch.cuda.synchronize()
    upd_gbs = P["X"].numel() * 96.0 / (u0.elapsed_time(u1) / 5 * 1e-3) * 1e-9

    if rank == 0:
        fl_launch = algorithmic_flops_per_unit(M, Din) * S * T * D           # one fused launch covers all units of the rank
        achieved = fl_launch / (fused_ms / max(fused_n, 1) * 1e-3) * 1e-12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fused_s = fused_ms / max(fused_n, 1) * 1e-3
        executed = executed_flops_per_unit(M, Din) * S * T * D / fused_s * 1e-12
        traffic = None
        try:      # DRAM bytes of the same launch from the committed ncu --set full capture (profiles/), when it is this workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json")))
            if tr.get("workload") == args.workload:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "GP transitions+grads/sec (S*T*D) per SG-HMC step", "value": value, "unit": "transitions+grads/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "T": T, "M": M, "D": D, "S_per_gpu": S, "kernel": "SquaredExponential",
                       "mode": "uncollapsed nll + all gradients + SG-HMC burn-in update of {X,U}",
                       "parallelism": "samples sharded, dp%d" % world,
                       "l2": "inputs larger than L2 (X and x-bar are %.0f MB each)" % (X_host.numel() * 8 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "transitions+grads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pipeline": "H2D of step k+1 (pinned host -> second device buffer pair, copy stream) overlaps step k; the first "
                                "copy is exposed; the host waits for every step's nll"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s", "frac": achieved / fp64_peak_tflops,
                         "traffic": traffic, "kernel": "ffvd::fused_kernel<SE,RB,NGW,UNCOLLAPSED>",
                         "convention": "achieved = SURVEY 8(d) algorithmic flops (dense 6M^2+9M*Din+8M per unit, what the reference's "
                                       "dense ops would issue) / fused-kernel time; the kernel skips the zero half of both triangular "
                                       "operands and the upper half of the SYRK, so frac can exceed 1 -- 'executed' is the honest "
                                       "tensor-pipe figure",
                         "executed": {"tflops": executed, "frac_of_peak": executed / fp64_peak_tflops,
                                      "flops_per_unit": executed_flops_per_unit(M, Din)},
                         "algorithmic_bytes_per_unit": 8.0 * (2 + (Din + 1 + 1) / D),
                         "peak_source": "measured live: cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json holds no FP64 figure); DMMA issue peak 37.15 TF (tools/probe)",
                         "algorithmic_flops_per_unit": algorithmic_flops_per_unit(M, Din), "fused_ms_per_launch": fused_ms / max(fused_n, 1),
                         "fused_share_of_step": (fused_ms / max(fused_n, 1)) / ms_per_step,
                         "fused_ms_per_rank": per_rank_fused,
                         "hbm_gbs_measured": peaks.get("hbm_gbs")},
            "roofline_update": {"bound": "hbm", "achieved": upd_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                "frac": (upd_gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                                "kernel": "ffvd::sghmc_kernel<burn_in> on X", "bytes_per_element": 96,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"},
            "clocks": clocks,
            "nll_mean": nll_check,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = time_cpu_oracle(cfg, 2, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
