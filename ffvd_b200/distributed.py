"""Multi-GPU plumbing: one process per GPU, samples (or whole chains) sharded across ranks with
no data-path collective; a single all-reduce of the packed shared-parameter gradients and the
scalar objective per evaluation.  `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests) is the transport; x-bar never leaves its owner.

Replicated parameters (Z, U, kernel hyper-parameters, logQ, C, d, logR) are updated redundantly on every rank from the
all-reduced gradient.  When such a parameter is SG-HMC sampled its injected noise must therefore be IDENTICAL on every
rank (draw it from a rank-independent seed, or broadcast it from rank 0 -- `shared_noise`); only the noise of the sharded
X is per rank.  Otherwise the replicas drift apart after the first update and the ranks evaluate different models.

When there are fewer trajectories than GPUs (one long chain), `time_block` / `evaluate_time_sharded` shard T instead
(SURVEY 8e): contiguous blocks of transitions with a one-row halo, uncollapsed form."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

SHARED = ("g_Z", "g_U", "g_logv", "g_logl", "g_logQ", "g_C", "g_d", "g_logR")


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def round_robin(n: int, rank: int, world: int) -> List[int]:
    """Whole independent chains dealt round-robin (BASELINE config 4); no collective needed."""
    return list(range(rank, n, world))


def shared_noise(shape, step: int, seed: int = 0, device=None):
    """N(0,1) noise for a REPLICATED parameter: a function of (seed, step) only, so every rank draws the same tensor."""
    import torch
    g = torch.Generator(device=device if device is not None else "cpu").manual_seed((int(seed) * 1000003 + int(step)) & 0x7FFFFFFFFFFFFFFF)
    return torch.randn(tuple(shape), dtype=torch.float64, device=device, generator=g)


def pack(outputs: Dict[str, object], names: Sequence[str] = SHARED):
    """Flatten the shared-parameter gradients (and nothing per-sample) into one buffer."""
    import torch
    present = [n for n in names if outputs.get(n) is not None]
    flat = torch.cat([outputs[n].reshape(-1) for n in present])
    return flat, present


def unpack(flat, outputs: Dict[str, object], present: Sequence[str]):
    o = 0
    for n in present:
        k = outputs[n].numel()
        outputs[n].copy_(flat[o:o + k].view_as(outputs[n]))
        o += k


def init_native_comm(ctx, group=None):
    """Create the context's own NCCL communicator (`ffvd_comm_init`, the collective behind the C ABI): rank 0 makes the
    128-byte unique id, `torch.distributed` (any backend) only ships it.  Afterwards `ctx.allreduce_shared(outputs)` packs,
    reduces and unpacks in three launches on the context's stream, with no torch op on the path."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(box[0], rank, world)
    return rank, world


def allreduce_shared(outputs: Dict[str, object], group=None, names: Sequence[str] = SHARED, ctx=None):
    """Sum the shared-parameter gradients over ranks in ONE collective (packed buffer).  With `ctx` (a context whose
    native communicator was set up by `init_native_comm`) the C-ABI collective is used; otherwise torch.distributed."""
    import torch.distributed as dist
    if ctx is not None and ctx.comm_info()[1] > 1:
        return ctx.allreduce_shared(outputs)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return outputs
    flat, present = pack(outputs, names)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    unpack(flat, outputs, present)
    return outputs


# ---- time sharding of ONE trajectory (S < number of GPUs) -----------------------------------------------------------
def time_block(problem: Dict[str, object], rank: int, world: int) -> Tuple[Dict[str, object], int, int]:
    """The block of transitions [a, b) owned by `rank`, as an ordinary problem: X rows a..b (b - a + 1 rows: the last
    one is the halo row x_b owned by the next rank), Y and ctrl rows a..b-1; everything else shared.
    Works on NumPy arrays and torch tensors (views, no copies of the parameters)."""
    X = problem["X"]
    if X.ndim != 2:
        raise ValueError("time sharding is for one trajectory: X must be (T+1, D)")
    T = X.shape[0] - 1
    a, b = shard_range(T, rank, world)
    blk = dict(problem)
    blk["X"] = X[a:b + 1]
    blk["Y"] = problem["Y"][a:b]
    if problem.get("ctrl") is not None:
        blk["ctrl"] = problem["ctrl"][a:b]
    return blk, a, b


def time_block_flags(rank: int) -> int:
    """Block 0 carries the shared-parameter priors and the x_0 prior; the other blocks drop both."""
    from . import _capi
    return 0 if rank == 0 else (_capi.FLAG_NO_SHARED_PRIORS | _capi.FLAG_NO_X0_PRIOR)


def evaluate_time_sharded(evaluate, problem: Dict[str, object], outputs: Dict[str, object], rank: int, world: int, group=None):
    """nll and gradients of ONE trajectory whose T transitions are split over `world` ranks (uncollapsed form).

    `evaluate(block_problem, block_outputs, extra_flags)` runs the ordinary single-GPU evaluation on a block (the CUDA
    path: `ctx.nll_grads(kind, False, blk, out, flags=base | extra_flags)`).  `outputs` holds this rank's tensors:
    `g_X` with the block's b - a + 1 rows, `nll` (1,), `terms` (1,6) and the shared-parameter gradients.  After the
    call: nll / terms / shared gradients are the FULL-trajectory values on every rank (one packed all-reduce);
    g_X rows a..b-1 (plus row b on the last rank) are final on their owner -- the halo row's contribution was sent
    to the next rank and the first row received the previous rank's (one all-gather of D doubles per rank).
    Every block evaluates with its own T_b in the 1/T normalisation (dgp_model.py:264-297), so everything is
    rescaled by T_b / T before the reduction."""
    import torch
    import torch.distributed as dist
    T = problem["X"].shape[0] - 1
    blk, a, b = time_block(problem, rank, world)
    evaluate(blk, outputs, time_block_flags(rank))
    scale = float(b - a) / float(T)
    for k, v in outputs.items():
        if v is not None:
            v.mul_(scale)
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return outputs
    names = ("nll", "terms") + SHARED
    flat, present = pack(outputs, names)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    unpack(flat, outputs, present)
    gX = outputs.get("g_X")
    if gX is not None:
        last = gX[-1].clone().contiguous()                 # contribution of this block to the halo row x_b
        rows = [torch.empty_like(last) for _ in range(world)]
        dist.all_gather(rows, last, group=group)
        if rank > 0:
            gX[0].add_(rows[rank - 1])                     # row a = halo row of the previous block
    return outputs


def evaluate_time_sharded_collapsed(ctx, kind: int, problem: Dict[str, object], outputs: Dict[str, object], rank: int, world: int,
                                    flags: int = 1, jitter: float = 1e-5, group=None, stats_transport: str = "native"):
    """The COLLAPSED bound (conditionals_multi_output.py:230-257, the CLI-default case 4) of ONE trajectory whose T
    transitions are split over `world` ranks.  H = F^T F / Q + I needs the statistics of ALL transitions, so the
    evaluation is split in two (include/ffvd_b200.h):
      pass 1 over the own block (FLAG_COLLAPSED_P1_ONLY) -> all-reduce of S = F^T F and b = F^T delta
      (`stats_transport` "native": the context's NCCL communicator; "torch": torch.distributed on copies)
      -> resume (FLAG_COLLAPSED_RESUME); every rank but 0 sets FLAG_NO_REPLICATED so that what depends on the global
      statistics only (log det H, the quadratic form, their log Q gradient, the Cholesky backward of G) is counted once.
    Then, as for the uncollapsed blocks: rescale by T_block / T, one packed all-reduce of nll / terms / shared gradients,
    one-row halo for g_X.  `outputs` as in `evaluate_time_sharded` (g_X with the block's b - a + 1 rows)."""
    import torch
    import torch.distributed as dist
    from . import _capi
    T = problem["X"].shape[0] - 1
    blk, a, b = time_block(problem, rank, world)
    blk = {k: (v.contiguous() if v is not None and hasattr(v, "contiguous") else v) for k, v in blk.items()}
    fl = int(flags) | time_block_flags(rank)
    ctx.nll_grads(kind, True, blk, outputs, flags=fl | _capi.FLAG_COLLAPSED_P1_ONLY, jitter=jitter)
    multi = world > 1 and dist.is_available() and dist.is_initialized()
    if multi:
        if stats_transport == "native" and ctx.comm_info()[1] > 1:
            ctx.collapsed_stats_allreduce()
        else:
            nb, Mp = ctx.collapsed_stats_shape()
            X = problem["X"]
            S = torch.empty((nb, Mp, Mp), dtype=torch.float64, device=X.device)
            bv = torch.empty((nb, Mp), dtype=torch.float64, device=X.device)
            ctx.collapsed_stats_get(S, bv)
            dist.all_reduce(S, group=group); dist.all_reduce(bv, group=group)
            ctx.collapsed_stats_set(S, bv)
    ctx.nll_grads(kind, True, blk, outputs, flags=fl | _capi.FLAG_COLLAPSED_RESUME | (_capi.FLAG_NO_REPLICATED if rank > 0 else 0),
                  jitter=jitter)
    scale = float(b - a) / float(T)
    for k, v in outputs.items():
        if v is not None:
            v.mul_(scale)
    if not multi:
        return outputs
    names = ("nll", "terms") + SHARED
    flat, present = pack(outputs, names)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    unpack(flat, outputs, present)
    gX = outputs.get("g_X")
    if gX is not None:
        last = gX[-1].clone().contiguous()
        rows = [torch.empty_like(last) for _ in range(world)]
        dist.all_gather(rows, last, group=group)
        if rank > 0:
            gX[0].add_(rows[rank - 1])
    return outputs


def exchange_halo_row(X_block, rank: int, world: int, group=None):
    """After the owners updated their rows (SG-HMC / Adam), refresh every block's halo row x_b from the next rank's
    first row (one all-gather of D doubles per rank)."""
    import torch
    import torch.distributed as dist
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return X_block
    first = X_block[0].clone().contiguous()
    rows = [torch.empty_like(first) for _ in range(world)]
    dist.all_gather(rows, first, group=group)
    if rank + 1 < world:
        X_block[-1].copy_(rows[rank + 1])
    return X_block
