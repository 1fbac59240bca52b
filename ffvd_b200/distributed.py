"""Multi-GPU plumbing: one process per GPU, samples (or whole chains) sharded across ranks with
no data-path collective; a single all-reduce of the packed shared-parameter gradients and the
scalar objective per evaluation.  `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests) is the transport; x-bar never leaves its owner."""
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


def allreduce_shared(outputs: Dict[str, object], group=None, names: Sequence[str] = SHARED):
    """Sum the shared-parameter gradients over ranks in ONE collective (packed buffer)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return outputs
    flat, present = pack(outputs, names)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    unpack(flat, outputs, present)
    return outputs
