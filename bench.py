#!/usr/bin/env python
"""bench.py -- FFVD GPSSM transitions+grads/sec per SG-HMC step on B200 (BASELINE.json metric).

A "step" is ONE evaluation of the GPSSM nll and all its gradients (uncollapsed q(u), SE kernel,
float64) on synthetic data of the named shape, the (N>1) all-reduce of the shared-parameter
gradients, and ONE adaptive SG-HMC burn-in update of the sampled set {X, U} (case-7 style)
with injected noise.  Units = S*T*D transitions per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c3small] [--extras auto|none|all]

N>1 is launched by the driver under torchrun (one rank per GPU, samples sharded S/rank = 64 each,
weak scaling).  --impl reference times the CPU restatement of the reference graph (oracle/) on a
bounded sample of the same workload on the box's host cores (TensorFlow itself is not installable).

Besides the headline numbers the JSON line carries
  parity : the CUDA path evaluated on the SAME inputs as the CPU oracle leg (T=20k, S=1 of the workload's M, D) and
           compared tensor by tensor (max-norm relative error; the run exits non-zero above 1e-9); at N>1 the
           all-reduced shared gradients of a small sample-sharded problem against rank 0's single-GPU evaluation.
  extra  : BASELINE.json's other named configurations measured in the same run -- c4_95chains (configs[3]: the 95
           bundled warm starts dealt round-robin over the ranks) and, at N>=2, c3_strong (configs[2] with its S=64 split
           over the ranks), time_sharded_chain (one trajectory split over time, uncollapsed and collapsed) and c5
           (configs[4]: T=1M, M=512, D=16, 32 trajectories per rank = the full S=256 job at N=8).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
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
    # same shard with T cut to 100k so that a step takes seconds instead of ~17 s (not the headline line: --workload only)
    "c5": dict(T=1_000_000, M=512, D=16, S=32, name="synthetic GPSSM T=1M M=512 D=16 S=256/8 per GPU SE float64 (BASELINE configs[4] shard)"),
    "c5short": dict(T=100_000, M=512, D=16, S=32, name="synthetic GPSSM T=100k M=512 D=16 S=32 per GPU SE float64 (configs[4] shape, T cut 10x)"),
}
CPU_SAMPLE_T = 20_000          # bounded CPU sample: T=20k, S=1 of the same M, D
PARITY_TOL = 1e-9              # north star: <= 1e-9 relative in float64 (relative to each tensor's max-norm)
PARAMS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR")
KERNEL_TAG = "r02b"            # version of the fused kernel the committed ncu traffic figure belongs to (profiles/fused_traffic.json)


def algorithmic_flops_per_unit(M, Din):
    return 6 * M * M + 9 * M * Din + 8 * M          # SURVEY 8(d), dense convention


def padded_M(M):
    return next(s for s in (64, 128, 256, 384, 512, 768, 1024, 1536, 2048) if M <= s)


def executed_flops_per_unit(M, Din):
    """Flops the fused kernel actually issues per (s,t,d) unit: both triangular contractions skip
    the zero half of L^{-1} (16-column groups), the SYRK forms only the 8x8 blocks on or below the diagonal of its 32x32
    tiles, plus the two thin products of the W back-propagation and the FP64 scalar work of the K tile."""
    Mp = padded_M(M)
    G, nt, nbx = Mp // 16, Mp // 32, (Din + 1 + 7) // 8
    tri = 128 * G * (G + 1)                                   # MACs per row of one triangular contraction
    syrk = 64 * (16 * (nt * (nt - 1) // 2) + 10 * nt)         # MACs per row of S += a a^T
    thin = 2 * Mp * 8 * nbx                                   # W^T [X,1] and W [Z,1]
    scalar = (Din + 25) * Mp                                  # r^2, exp, statistics
    return 2 * (2 * tri + syrk + thin) + scalar


def shared_params(M, D):
    """Parameters shared by every sample / rank of a synthetic workload (SURVEY 8(d) recipe)."""
    Din = D + 1
    shared = np.random.default_rng(20230209)
    Z = shared.standard_normal((M, Din)) * 1.5
    logl = np.log(shared.uniform(1.0, 4.0, (D, Din)))
    logv = np.log(shared.uniform(0.05, 0.8, D))
    logQ = 2.0 * np.log(shared.uniform(0.2, 0.8, D))
    logR = np.log(np.full((1, 1), 0.4))
    C = shared.standard_normal((D, 1)) * 0.3
    d = np.zeros(1)
    U = shared.standard_normal((M, D))
    return dict(Z=Z, U=U, logv=logv, logl=logl, logQ=logQ, C=C, d=d, logR=logR)


def make_host_data(T, M, D, S, seed):
    """SURVEY 8(d) synthetic recipe (AR(1) trajectories, N(0,1.5^2) inducing inputs, ...), on the host."""
    from scipy.signal import lfilter
    rng = np.random.default_rng(seed)
    ctrl = rng.standard_normal((T, 1))
    eps = rng.standard_normal((S, T + 1, D))
    eps[:, 1:] *= np.sqrt(1 - 0.95 ** 2)
    X = lfilter([1.0], [1.0, -0.95], eps, axis=1)             # x_{t+1} = 0.95 x_t + e_t, x_0 ~ N(0,1)
    sp = shared_params(M, D)                                  # shared parameters identical on every rank
    Y = X[0, 1:] @ sp["C"] + sp["d"] + 0.4 * rng.standard_normal((T, 1))
    return dict(X=np.ascontiguousarray(X), Y=Y, ctrl=ctrl, **sp)


def ar1_filter_device(eps, rho=0.95, block=256):
    """x_t = rho x_{t-1} + eps_t along dim 1 of eps (S, N, D), on the device: a blocked scan -- every block of `block`
    steps is one small lower-triangular Toeplitz product, the carry between blocks is a short sequential loop."""
    import torch
    S, N, D = eps.shape
    nb = (N + block - 1) // block
    pad = nb * block - N
    if pad:
        eps = torch.cat((eps, torch.zeros((S, pad, D), dtype=eps.dtype, device=eps.device)), dim=1)
    i = torch.arange(block, device=eps.device)
    A = torch.where(i[:, None] >= i[None, :], rho ** (i[:, None] - i[None, :]).clamp(min=0).to(eps.dtype),
                    torch.zeros((), dtype=eps.dtype, device=eps.device))
    x = torch.matmul(A, eps.view(S, nb, block, D))            # (S, nb, block, D): within-block responses
    del eps
    powv = (rho ** (i + 1).to(x.dtype))[None, :, None]        # rho^(i+1)
    carry = torch.zeros((S, 1, D), dtype=x.dtype, device=x.device)
    for b in range(nb):
        xb = x[:, b]
        xb.add_(powv * carry)
        carry = xb[:, -1:, :].clone()
    return x.view(S, nb * block, D)[:, :N].contiguous()


def make_device_data(T, M, D, S, seed, dev):
    """The same recipe generated on the device (for shapes whose X is too large to ship from the host --
    configs[4] is 4.1 GB of X per GPU).  The trajectories use torch's device generator, so they are not the host
    recipe's numbers; the AR(1) filter itself is validated against scipy in tests/test_gpu_parity.py."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    f64 = torch.float64
    ctrl = torch.randn((T, 1), dtype=f64, device=dev, generator=g)
    eps = torch.randn((S, T + 1, D), dtype=f64, device=dev, generator=g)
    eps[:, 1:] *= float(np.sqrt(1 - 0.95 ** 2))
    X = ar1_filter_device(eps)
    del eps
    sp = {k: torch.as_tensor(v, dtype=f64, device=dev).contiguous() for k, v in shared_params(M, D).items()}
    Y = X[0, 1:] @ sp["C"] + sp["d"] + 0.4 * torch.randn((T, 1), dtype=f64, device=dev, generator=g)
    return dict(X=X, Y=Y.contiguous(), ctrl=ctrl, **sp)


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


def cpu_sample_problem(cfg):
    """The bounded CPU sample of a workload: T = 20k, S = 1 of the same M, D (seed 1) -- shared by the CPU oracle leg and
    the parity block, so the CUDA path is checked on exactly the inputs the oracle is timed on."""
    from oracle.ffvd_oracle import Problem
    T = min(CPU_SAMPLE_T, cfg["T"])
    h = make_host_data(T, cfg["M"], cfg["D"], 1, seed=1)
    prob = Problem(X=h["X"][0], Z=h["Z"], U=h["U"], logv=h["logv"], logl=h["logl"], logQ=h["logQ"], C=h["C"], d=h["d"],
                   logR=h["logR"], Y=h["Y"], ctrl=h["ctrl"])
    return prob, h, T


def time_cpu_oracle(cfg, steps, warmup):
    """CPU restatement of the reference graph (oracle/, torch CPU float64 + autograd) on a bounded sample.
    Returns the timing record and the oracle's nll / gradients of the FIRST evaluation (the parity reference)."""
    import torch
    from oracle import ffvd_oracle as O
    # all the host threads available (torchrun exports OMP_NUM_THREADS=1 to its workers: undo that for the CPU arm)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    prob, _, T = cpu_sample_problem(cfg)
    rng = np.random.default_rng(0)
    st = {n: [np.ones_like(getattr(prob, n)), np.ones_like(getattr(prob, n)), np.ones_like(getattr(prob, n)), np.zeros_like(getattr(prob, n))]
          for n in ("X", "U")}
    times, first = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        res = O.nll_and_grads(prob, collapsed=False)
        if first is None:
            first = {k: np.array(v, dtype=np.float64, copy=True) for k, v in res.items()}
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
    med, best = float(np.median(times)), float(np.min(times))
    rec = dict(value=units / med, unit="transitions+grads/s", cores=torch.get_num_threads(), kind="port",
               sample="oracle (torch CPU float64 restatement of the TF graph + autograd) on T=%d, S=1, M=%d, D=%d of the same synthetic "
                      "workload, %d timed steps, median %.3f s/step (min %.3f, max %.3f); host has %d logical cpus"
                      % (T, cfg["M"], cfg["D"], len(times), med, best, float(np.max(times)), os.cpu_count()),
               ms_per_step=1e3 * med, value_best=units / best, steps=len(times))
    return rec, first


def relerr(ref, got):
    ref = np.asarray(ref, dtype=np.float64); got = np.asarray(got, dtype=np.float64).reshape(ref.shape)
    if not np.all(np.isfinite(got)):
        return float("inf")
    return float(np.max(np.abs(ref - got)) / max(float(np.max(np.abs(ref))), 1e-300)) if ref.size else 0.0


def alloc_outputs(P, S, dev):
    import torch
    out = {"nll": torch.empty(S, dtype=torch.float64, device=dev), "terms": torch.empty(S, 6, dtype=torch.float64, device=dev)}
    for k in PARAMS:
        out["g_" + k] = torch.empty_like(P[k])
    return out


def cuda_parity_vs_oracle(ctx, cfg, dev, oracle_first):
    """Evaluate the CUDA path (through the C ABI) on the inputs of the CPU oracle leg and compare every output."""
    import torch
    import ffvd_b200
    _, h, T = cpu_sample_problem(cfg)
    P = {k: torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for k, v in h.items()}
    P["X"] = P["X"][0].contiguous()
    out = alloc_outputs(P, 1, dev)
    ctx.nll_grads(ffvd_b200.KERNEL_SE, False, P, out, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL)
    torch.cuda.synchronize()
    errs = {k: relerr(oracle_first[k], v.cpu().numpy()) for k, v in out.items() if k in oracle_first}
    return {"inputs": "T=%d S=1 M=%d D=%d (the cpu_baseline sample)" % (T, cfg["M"], cfg["D"]), "tol": PARITY_TOL,
            "max_rel_err": errs, "worst": max(errs.values()), "ok": bool(max(errs.values()) <= PARITY_TOL)}


def multi_gpu_parity(ctx, dev, rank, world):
    """A small sample-sharded problem (2 trajectories per rank): all-reduced shared gradients and per-sample nll against
    rank 0's single-GPU evaluation of all 2*world trajectories."""
    import torch
    import torch.distributed as dist
    import ffvd_b200
    from ffvd_b200 import distributed as fd
    T, M, D, Sp = 2048, 256, 8, 2
    h = make_host_data(T, M, D, Sp * world, seed=77)          # same data on every rank, each takes its slice
    full = {k: torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for k, v in h.items()}
    mine = dict(full); mine["X"] = full["X"][rank * Sp:(rank + 1) * Sp].contiguous()
    out = alloc_outputs(mine, Sp, dev)
    ctx.nll_grads(ffvd_b200.KERNEL_SE, False, mine, out, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL)
    ctx.allreduce_shared(out)
    nll_all = [torch.empty(Sp, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(nll_all, out["nll"])
    res = None
    if rank == 0:
        ref = alloc_outputs(full, Sp * world, dev)
        ctx.nll_grads(ffvd_b200.KERNEL_SE, False, full, ref, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL)
        torch.cuda.synchronize()
        errs = {k: relerr(ref[k].cpu().numpy(), out[k].cpu().numpy()) for k in fd.SHARED}
        errs["nll"] = relerr(ref["nll"].cpu().numpy(), torch.cat(nll_all).cpu().numpy())
        errs["g_X(rank0 shard)"] = relerr(ref["g_X"][:Sp].cpu().numpy(), out["g_X"].cpu().numpy())
        res = {"inputs": "T=%d M=%d D=%d, %d trajectories per rank x %d ranks vs rank 0 alone" % (T, M, D, Sp, world),
               "max_rel_err": errs, "worst": max(errs.values()), "tol": PARITY_TOL, "ok": bool(max(errs.values()) <= PARITY_TOL)}
    dist.barrier()
    return res


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    steps, warmup = max(5, min(args.steps, 8)), max(1, min(args.warmup, 2))
    cb, _ = time_cpu_oracle(cfg, steps, warmup)
    line = {"impl": "reference", "metric": "GP transitions+grads/sec (S*T*D) per SG-HMC step", "value": cb["value"],
            "unit": "transitions+grads/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "sample": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "value_best")},
            "e2e": {"value": cb["value"], "unit": "transitions+grads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was re-routed to stderr."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


class StepRunner:
    """One rank's resident state for a synthetic workload and the `step` the bench times."""

    def __init__(self, ctx, dev, P, T, S, world, rank, host_noise=False):
        import torch
        import ffvd_b200
        self.ctx, self.dev, self.P, self.T, self.S, self.world = ctx, dev, P, T, S, world
        self.out = alloc_outputs(P, S, dev)
        # the X noise is per rank (X is sharded); the noise of the REPLICATED parameter U must be identical on every rank,
        # or the replicas drift apart after the first update (see ffvd_b200/distributed.py)
        gx = torch.Generator(device=dev).manual_seed(4321 + rank)
        gu = torch.Generator(device=dev).manual_seed(987)
        self.noise_host = None
        if host_noise:
            torch.manual_seed(4321 + rank)
            self.noise_host = torch.randn(P["X"].shape, dtype=torch.float64).pin_memory()
            self.noise_X = self.noise_host.to(dev)
        else:
            self.noise_X = torch.randn(P["X"].shape, dtype=torch.float64, device=dev, generator=gx)
        self.noise_U = torch.randn(P["U"].shape, dtype=torch.float64, device=dev, generator=gu)
        self.state = {n: dict(xi=torch.ones_like(P[n]), g=torch.ones_like(P[n]), g2=torch.ones_like(P[n]), p=torch.zeros_like(P[n]))
                      for n in ("X", "U")}
        self.U0 = P["U"].clone()
        self.flags = ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC
        self.kind = ffvd_b200.KERNEL_SE

    def step(self, Pd=None, nzX=None):
        from ffvd_b200 import distributed as fd
        Pd = self.P if Pd is None else Pd
        nzX = self.noise_X if nzX is None else nzX
        self.ctx.nll_grads(self.kind, False, Pd, self.out, flags=self.flags)
        if self.world > 1:
            # one packed NCCL all-reduce of the Z/U/hyper-parameter gradients, through the library's own communicator
            # (ffvd_allreduce_shared: pack, ncclAllReduce, unpack on the context's stream)
            self.ctx.allreduce_shared(self.out)
        for n, nz in (("X", nzX), ("U", self.noise_U)):
            st = self.state[n]
            self.ctx.sghmc_update(Pd[n], self.out["g_" + n], nz, st["xi"], st["g"], st["g2"], st["p"], 0.01, 0.05, float(self.T + 1), True)

    def reset(self, X0=None):
        if X0 is not None:
            self.P["X"].copy_(X0)
        self.P["U"].copy_(self.U0)
        for n in ("X", "U"):
            st = self.state[n]
            st["xi"].fill_(1.0); st["g"].fill_(1.0); st["g2"].fill_(1.0); st["p"].zero_()


def timed_steps(runner, K, world, barrier):
    """K steps bracketed by barrier + synchronize, CUDA events, max over ranks -> (ms per step, fused ms per launch)."""
    import torch
    import torch.distributed as dist
    barrier()
    runner.ctx.fused_time(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        runner.step()
    e1.record()
    barrier()
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=runner.dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    fused_ms, fused_n = runner.ctx.fused_time(reset=True)
    return float(tmax.item()) / K, fused_ms / max(fused_n, 1)


def gather_floats(x, dev, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return [float(x)]
    tl = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(tl, torch.tensor([float(x)], dtype=torch.float64, device=dev))
    return [float(t.item()) for t in tl]


def extra_synthetic(ctx, dev, rank, world, barrier, name, T, M, D, S_rank, warm, K, fp64_peak, note):
    """One more synthetic configuration in the same run (device-generated data): ms/step, units/s, roofline fractions."""
    import torch
    P = make_device_data(T, M, D, S_rank, seed=5000 + rank, dev=dev)
    runner = StepRunner(ctx, dev, P, T, S_rank, world, rank)
    for _ in range(warm):
        runner.step()
    ms, fused = timed_steps(runner, K, world, barrier)
    nll = float(runner.out["nll"].mean().item())
    per_rank = gather_floats(fused, dev, world)
    Din = D + 1
    units = S_rank * T * D * world
    ex = executed_flops_per_unit(M, Din) * S_rank * T * D / (max(per_rank) * 1e-3) * 1e-12
    al = algorithmic_flops_per_unit(M, Din) * S_rank * T * D / (max(per_rank) * 1e-3) * 1e-12
    del runner, P
    torch.cuda.empty_cache()
    return {"config": name, "T": T, "M": M, "D": D, "S_per_gpu": S_rank, "S_total": S_rank * world, "n_gpus": world, "warmup": warm, "steps": K,
            "ms_per_step": ms, "value": units / (ms * 1e-3), "unit": "transitions+grads/s", "fused_ms_per_rank": per_rank,
            "executed_tflops_per_gpu": ex, "executed_frac_of_peak": ex / fp64_peak, "algorithmic_tflops_per_gpu": al,
            "algorithmic_frac": al / fp64_peak, "nll_mean": nll, "finite": bool(np.isfinite(nll)),
            "data": "synthetic, generated on the device", "note": note}


def extra_c4(ctx, dev, rank, world, barrier, K):
    """BASELINE configs[3]: the 95 bundled warm starts (6 datasets) as independent chains dealt round-robin over the
    ranks (`distributed.round_robin`), one batched call per rank, no collective.  chains*T*D per second."""
    import torch
    import torch.distributed as dist
    import ffvd_b200
    from ffvd_b200 import distributed as fd
    from ffvd_b200.datasets import load_packed_problems
    probs = load_packed_problems(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    mine = fd.round_robin(len(probs), rank, world)
    res = {}
    for collapsed in (False, True):
        Ps, Os, units = [], [], 0
        for i in mine:
            P = {k: (None if v is None else torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous()) for k, v in probs[i].items() if k != "name"}
            Ps.append(P); Os.append(alloc_outputs(P, 1, dev)); units += (P["X"].shape[0] - 1) * P["X"].shape[1]
        call = ctx.prepare_nll_grads(ffvd_b200.KERNEL_SE, collapsed, Ps, Os, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL | ffvd_b200.FLAG_ASYNC)
        for _ in range(3):
            call.run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            call.run()
        e1.record()
        barrier()
        tmax = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
        tot = torch.tensor([float(units)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        finite = all(bool(torch.isfinite(o["nll"]).all().item()) for o in Os)
        res["collapsed" if collapsed else "uncollapsed"] = {
            "ms_per_evaluation_of_all_chains": float(tmax.item()), "chains_T_D_per_s": float(tot.item()) / (float(tmax.item()) * 1e-3),
            "chains_on_rank0": len(mine), "finite": finite}
        call.close()
    res.update({"config": "BASELINE configs[3]: 95 warm starts x 6 bundled datasets, chains dealt round-robin over %d GPU(s), one batched call "
                          "per GPU, nll + all gradients per evaluation, no collective" % world, "chains": len(probs), "evaluations_timed": K})
    return res


def extra_time_sharded(ctx, dev, rank, world, barrier, K):
    """ONE long trajectory (S = 1 < number of GPUs) sharded over TIME: contiguous blocks of transitions with a
    one-row halo; uncollapsed (one packed all-reduce + halo gather) and collapsed (pass 1, all-reduce of F^T F / F^T delta
    through the library's communicator, resume).  Evaluations only (nll + all gradients), device-generated data identical on
    every rank (same seed), each rank takes its block."""
    import torch
    import torch.distributed as dist
    import ffvd_b200
    from ffvd_b200 import distributed as fd
    T, M, D = 800_000, 256, 8
    P = make_device_data(T, M, D, 1, seed=31337, dev=dev)
    P["X"] = P["X"][0].contiguous()
    a, b = fd.shard_range(T, rank, world)
    out = {"nll": torch.empty(1, dtype=torch.float64, device=dev), "terms": torch.empty(1, 6, dtype=torch.float64, device=dev),
           "g_X": torch.empty((b - a + 1, D), dtype=torch.float64, device=dev)}
    for k in PARAMS[1:]:
        out["g_" + k] = torch.empty_like(P[k])
    FL = ffvd_b200.FLAG_PRIOR_Z_NORMAL

    def ev(blk_p, blk_o, extra):
        bp = {k: (v.contiguous() if v is not None else None) for k, v in blk_p.items()}
        ctx.nll_grads(ffvd_b200.KERNEL_SE, False, bp, blk_o, flags=FL | extra | ffvd_b200.FLAG_ASYNC)

    res = {}
    for tag, fn in (("uncollapsed", lambda: fd.evaluate_time_sharded(ev, P, out, rank, world)),
                    ("collapsed", lambda: fd.evaluate_time_sharded_collapsed(ctx, ffvd_b200.KERNEL_SE, P, out, rank, world, flags=FL | ffvd_b200.FLAG_ASYNC))):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        barrier()
        tmax = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        res[tag] = {"ms_per_evaluation": float(tmax.item()), "T_D_per_s": T * D / (float(tmax.item()) * 1e-3), "nll": float(out["nll"][0].item())}
    res["config"] = "one trajectory T=%d M=%d D=%d (S=1) split into %d time blocks, nll + all gradients per evaluation" % (T, M, D, world)
    del P, out
    torch.cuda.empty_cache()
    return res


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
    ap.add_argument("--extras", default="auto", choices=["auto", "none", "all"],
                    help="auto: c4_95chains always; c3_strong and c5 (T=1M) when N >= 2 and the workload is c3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, cfg, rank, world)

    import torch
    import torch.distributed as dist
    import ffvd_b200
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
    if world > 1:
        from ffvd_b200 import distributed as fd
        fd.init_native_comm(ctx)                    # NCCL communicator owned by the library (C ABI); torch only ships the id

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- data: host (pinned) copies for the e2e legs, device-resident copies for `value`
    h = make_host_data(T, M, D, S, seed=1000 + rank)
    P = {k: torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for k, v in h.items()}
    X_host = torch.as_tensor(h["X"]).pin_memory()
    runner = StepRunner(ctx, dev, P, T, S, world, rank, host_noise=True)
    out, noise_X, noise_host = runner.out, runner.noise_X, runner.noise_host
    X0 = P["X"].clone()
    step = runner.step

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
    runner.reset(X0)
    barrier()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    ms_per_step, fused_ms_launch = timed_steps(runner, K, world, barrier)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count - launches0
    per_rank_fused = gather_floats(fused_ms_launch, dev, world)
    units = S * T * D * world
    value = units / (ms_per_step * 1e-3)
    nll_check = float(out["nll"].mean().item())

    # ---- e2e (1): host buffers, H2D of the step's inputs and D2H of its result inside the timed region
    runner.reset(X0)
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

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(n)
        e1.record()
        barrier()
        tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return units / (float(tmax.item()) / n * 1e-3)

    run_e2e(2)
    e2e_value = timed(run_e2e, K)
    h2d = X_host.numel() * 8 + noise_host.numel() * 8
    d2h = S * 8
    Pb[1].pop("X"); Nb.pop()

    # ---- e2e (2): a chain whose trajectory LIVES ON THE HOST between steps: every step uploads the current X (and its
    #      noise), evaluates, updates, and downloads the updated X, which is what the next step uploads.  The X round trip
    #      cannot be hidden (step k+1 needs step k's result); only the noise upload of step k+1 overlaps step k.
    runner.reset(X0)
    Xh = X_host.clone().pin_memory()
    Nb2 = [noise_X, torch.empty_like(noise_X)]
    nready = [torch.cuda.Event(), torch.cuda.Event()]
    nfree = [torch.cuda.Event(), torch.cuda.Event()]

    def noise_prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(nfree[i])
            Nb2[i].copy_(noise_host, non_blocking=True)
            nready[i].record(copy_stream)

    def run_e2e_host_state(n):
        noise_prefetch(0)
        for k in range(n):
            i = k & 1
            P["X"].copy_(Xh, non_blocking=True)                 # H2D of the chain state
            main_stream.wait_event(nready[i])
            if k + 1 < n:
                noise_prefetch(1 - i)
            step(P, Nb2[i])
            nfree[i].record(main_stream)
            Xh.copy_(P["X"], non_blocking=True)                 # D2H of the updated state
            nll_host.copy_(out["nll"], non_blocking=True)
            main_stream.synchronize()

    run_e2e_host_state(2)
    e2e_hs_value = timed(run_e2e_host_state, K)
    del Nb2, Xh

    # ---- the second kernel of the step against its own roofline: the SG-HMC burn-in update of X is an HBM
    #      stream of 96 B / element (7 reads + 5 writes)
    runner.reset(X0)
    st = runner.state["X"]
    for _ in range(2):
        ctx.sghmc_update(P["X"], out["g_X"], noise_X, st["xi"], st["g"], st["g2"], st["p"], 0.01, 0.05, float(T + 1), True)
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(5):
        ctx.sghmc_update(P["X"], out["g_X"], noise_X, st["xi"], st["g"], st["g2"], st["p"], 0.01, 0.05, float(T + 1), True)
    u1.record(); torch.cuda.synchronize()
    upd_gbs = P["X"].numel() * 96.0 / (u0.elapsed_time(u1) / 5 * 1e-3) * 1e-9

    # ---- free the headline workload, then parity and the extra configurations
    del runner, step, out, noise_X, noise_host, X0, X_host, P, Pb, Nb, h, st
    torch.cuda.empty_cache()

    parity = {}
    cpu_rec = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_rec, oracle_first = time_cpu_oracle(cfg, 5, 1)
        parity["cuda_vs_oracle"] = cuda_parity_vs_oracle(ctx, cfg, dev, oracle_first)
    if world > 1:
        mg = multi_gpu_parity(ctx, dev, rank, world)
        if rank == 0:
            parity["multi_gpu_allreduce"] = mg

    extra = {}
    if args.extras != "none":
        try:
            extra["c4_95chains"] = extra_c4(ctx, dev, rank, world, barrier, 20)
        except Exception as e:                       # fixtures missing: say so instead of dropping the key silently
            extra["c4_95chains"] = {"error": repr(e)}
        want_big = (args.extras == "all") or (args.extras == "auto" and world >= 2 and args.workload == "c3")
        if want_big and world >= 2 and 64 % world == 0:
            extra["c3_strong"] = extra_synthetic(ctx, dev, rank, world, barrier, "BASELINE configs[2] strong scaling: S=64 split over the ranks",
                                                 100_000, 256, 8, 64 // world, 3, 5, fp64_peak_tflops,
                                                 "total work fixed (S=64): compare value with the N=1 headline line")
        if want_big and world >= 2:
            extra["time_sharded_chain"] = extra_time_sharded(ctx, dev, rank, world, barrier, 5)
        if want_big:
            extra["c5"] = extra_synthetic(ctx, dev, rank, world, barrier,
                                          "BASELINE configs[4]: T=1M M=512 D=16, 32 trajectories per GPU%s"
                                          % (" = the full S=256 job" if world == 8 else " (the full S=256 needs 8 GPUs)"),
                                          1_000_000, 512, 16, 32, 2, 2, fp64_peak_tflops,
                                          "full T; 2 warm-up + 2 timed steps (a step is ~17 s); weak scaling in S: per-GPU work is that of "
                                          "the 8-GPU job at every N")

    if rank == 0:
        fl_launch = algorithmic_flops_per_unit(M, Din) * S * T * D           # one fused launch covers all units of the rank
        fused_s = fused_ms_launch * 1e-3
        achieved = fl_launch / fused_s * 1e-12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        executed = executed_flops_per_unit(M, Din) * S * T * D / fused_s * 1e-12
        traffic, traffic_source = None, None
        try:      # DRAM bytes per launch from the committed ncu --set full capture -- only for this workload AND this kernel version
            tr = json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json")))
            if tr.get("workload") == args.workload and tr.get("kernel_tag") == KERNEL_TAG and world == 1:
                traffic, traffic_source = tr["dram_bytes_per_launch"], tr.get("source")
        except Exception:
            pass
        line = {
            "metric": "GP transitions+grads/sec (S*T*D) per SG-HMC step", "value": value, "unit": "transitions+grads/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "T": T, "M": M, "D": D, "S_per_gpu": S, "kernel": "SquaredExponential",
                       "mode": "uncollapsed nll + all gradients + SG-HMC burn-in update of {X,U}",
                       "parallelism": "samples sharded, dp%d" % world,
                       "l2": "inputs larger than L2 (X and x-bar are %.0f MB each)" % (S * (T + 1) * D * 8 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "transitions+grads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pipeline": "H2D of step k+1 (pinned host -> second device buffer pair, copy stream) overlaps step k; the first "
                                "copy is exposed; the host waits for every step's nll",
                    "host_resident_state": {"value": e2e_hs_value, "unit": "transitions+grads/s",
                                            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": S * (T + 1) * D * 8 + d2h,
                                            "what": "the chain's X lives in pinned host memory between steps: H2D of the current X + its noise, "
                                                    "evaluate, SG-HMC update, D2H of the updated X (the next step's input) and of nll, every step; "
                                                    "only the noise upload overlaps the previous step"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": executed, "peak": fp64_peak_tflops, "unit": "TFLOP/s", "frac": executed / fp64_peak_tflops,
                         "traffic": traffic, "traffic_source": traffic_source, "kernel": "ffvd::fused_kernel<SE,RB,NGW,UNCOLLAPSED>",
                         "convention": "achieved / frac = flops the kernel EXECUTES on the FP64 tensor pipe (it skips the zero half of both "
                                       "triangular operands and the upper half of the SYRK) / fused-kernel time, CUDA events; "
                                       "algorithmic_* = SURVEY 8(d)'s dense convention (6M^2+9M*Din+8M per unit, what the reference's dense "
                                       "ops would issue), which can exceed the peak",
                         "executed_flops_per_unit": executed_flops_per_unit(M, Din),
                         "algorithmic_tflops": achieved, "algorithmic_frac": achieved / fp64_peak_tflops,
                         "algorithmic_flops_per_unit": algorithmic_flops_per_unit(M, Din),
                         "algorithmic_bytes_per_unit": 8.0 * (2 + (Din + 1 + 1) / D),
                         "peak_source": "measured live: cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json holds no FP64 figure); DMMA issue peak 37.15 TF (tools/probe)",
                         "fused_ms_per_launch": fused_ms_launch,
                         "fused_share_of_step": fused_ms_launch / ms_per_step,
                         "fused_ms_per_rank": per_rank_fused,
                         "hbm_gbs_measured": peaks.get("hbm_gbs")},
            "roofline_update": {"bound": "hbm", "achieved": upd_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                "frac": (upd_gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                                "kernel": "ffvd::sghmc_kernel<burn_in> on X", "bytes_per_element": 96,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"},
            "clocks": clocks,
            "nll_mean": nll_check,
            "parity": parity,
            "extra": extra,
        }
        line["cpu_baseline"] = {k: cpu_rec[k] for k in ("value", "unit", "cores", "kind", "sample", "value_best")} if cpu_rec else None
        emit(line)
    ok = all(v.get("ok", True) for v in parity.values()) if rank == 0 else True
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write("bench.py: PARITY FAILURE %s\n" % json.dumps(parity))
        sys.exit(3)


if __name__ == "__main__":
    main()
