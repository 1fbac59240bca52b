"""Multi-GPU correctness check (run under torchrun on >= 2 GPUs of one box):
  (1) sample sharding: every rank evaluates its block of trajectories, ONE packed NCCL all-reduce of the shared-parameter
      gradients -> equals the single-GPU evaluation of all S trajectories;
  (2) time sharding of one trajectory (S < #GPUs): blocks of transitions with a one-row halo -> equals the single-GPU
      evaluation of the whole trajectory;
  (3) the same two through the library's OWN NCCL communicator (ffvd_comm_init / ffvd_allreduce_shared, C ABI);
  (4) the COLLAPSED bound under time sharding (pass 1 -> ffvd_collapsed_stats_allreduce -> resume).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_multigpu.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ffvd_b200
from ffvd_b200 import distributed as fd
from oracle import fixtures

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = ffvd_b200.Context(local, torch.cuda.current_stream(local).cuda_stream)
KEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")
GK = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR")


def to_dev(prob):
    return {k: torch.as_tensor(np.ascontiguousarray(getattr(prob, k)), dtype=torch.float64, device=dev) for k in KEYS}


def alloc(p):
    S = 1 if p["X"].dim() == 2 else p["X"].shape[0]
    o = {"nll": torch.zeros(S, dtype=torch.float64, device=dev), "terms": torch.zeros(S, 6, dtype=torch.float64, device=dev)}
    for k in GK:
        o["g_" + k] = torch.zeros_like(p[k])
    return o


def rel(a, b):
    return float((a - b).abs().max() / a.abs().max().clamp_min(1e-300))


worst = 0.0
parts = {}


# ---- (1) sample sharding
S = 2 * world + 1
full_p = to_dev(fixtures.synthetic_problem(T=333, M=150, D=3, S=S))
full_o = alloc(full_p)
ctx.nll_grads(0, False, full_p, full_o)
lo, hi = fd.shard_range(S, rank, world)
mine = dict(full_p); mine["X"] = full_p["X"][lo:hi].contiguous()
o = alloc(mine)
ctx.nll_grads(0, False, mine, o)
fd.allreduce_shared(o)
for k in fd.SHARED:
    worst = max(worst, rel(full_o[k], o[k]))
worst = max(worst, rel(full_o["g_X"][lo:hi], o["g_X"]), rel(full_o["nll"][lo:hi], o["nll"]))
parts["sample_sharding"] = worst
w1 = worst; worst = 0.0
# ---- (2) time sharding of one trajectory
one_p = to_dev(fixtures.synthetic_problem(T=1001, M=150, D=3, S=1))
one_o = alloc(one_p)
ctx.nll_grads(0, False, one_p, one_o)
blk, a, b = fd.time_block(one_p, rank, world)
out = {"nll": torch.zeros(1, dtype=torch.float64, device=dev), "terms": torch.zeros(1, 6, dtype=torch.float64, device=dev),
       "g_X": torch.zeros(b - a + 1, 3, dtype=torch.float64, device=dev)}
for k in GK[1:]:
    out["g_" + k] = torch.zeros_like(one_p[k])


def evaluate(blk_p, blk_o, extra):
    bp = {k: (v.contiguous() if v is not None else None) for k, v in blk_p.items()}
    ctx.nll_grads(0, False, bp, blk_o, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL | extra)


fd.evaluate_time_sharded(evaluate, one_p, out, rank, world)
for k in ("nll", "terms") + fd.SHARED:
    worst = max(worst, rel(one_o[k].reshape(-1), out[k].reshape(-1)))
rows = slice(0, b - a + (1 if rank == world - 1 else 0))
worst = max(worst, rel(one_o["g_X"][a:a + rows.stop], out["g_X"][rows]))
parts["time_sharding_uncollapsed"] = worst
w2 = worst; worst = 0.0
# ---- (3) the library's own communicator
fd.init_native_comm(ctx)
assert ctx.comm_info()[1] == world
o2 = alloc(mine)
ctx.nll_grads(0, False, mine, o2)
ctx.allreduce_shared(o2)
for k in fd.SHARED:
    worst = max(worst, rel(full_o[k], o2[k]))
parts["native_communicator"] = worst
w3 = worst; worst = 0.0
# ---- (4) collapsed bound, time sharded, statistics all-reduced by the native communicator and by torch
col_o = alloc(one_p)
ctx.nll_grads(0, True, one_p, col_o)
for transport in ("native", "torch"):
    outc = {"nll": torch.zeros(1, dtype=torch.float64, device=dev), "terms": torch.zeros(1, 6, dtype=torch.float64, device=dev),
            "g_X": torch.zeros(b - a + 1, 3, dtype=torch.float64, device=dev)}
    for k in GK[1:]:
        outc["g_" + k] = torch.zeros_like(one_p[k])
    fd.evaluate_time_sharded_collapsed(ctx, 0, one_p, outc, rank, world, flags=ffvd_b200.FLAG_PRIOR_Z_NORMAL, stats_transport=transport)
    for k in ("nll", "terms") + fd.SHARED:
        worst = max(worst, rel(col_o[k].reshape(-1), outc[k].reshape(-1)))
    worst = max(worst, rel(col_o["g_X"][a:a + rows.stop], outc["g_X"][rows]))
parts["time_sharding_collapsed"] = worst
names = sorted(parts)
w = torch.tensor([parts[n] for n in names], dtype=torch.float64, device=dev)
dist.all_reduce(w, op=dist.ReduceOp.MAX)
if rank == 0:
    TOL = 1e-9     # the north star's float64 tolerance; the sharded sums differ from the single-GPU ones by their summation order only
    for n, v in zip(names, w.tolist()):
        print("  %-28s %.3e" % (n, v))
    print("multi-GPU check on %d GPUs: worst max-norm relative deviation from the single-GPU evaluation %.3e  %s (tolerance %.0e)"
          % (world, float(w.max().item()), "OK" if float(w.max().item()) <= TOL else "FAIL", TOL))
dist.destroy_process_group()
