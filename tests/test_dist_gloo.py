"""CPU test of the N>1 path (world_size 2, gloo): samples are sharded over ranks, every rank
evaluates its shard, and ONE packed all-reduce of the shared-parameter gradients reproduces the
single-process result.  The per-rank evaluation is the oracle here (the device path is covered by
the GPU tests); what is under test is the sharding + packing + collective host logic."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import copy
    from ffvd_b200 import distributed
    from oracle import ffvd_oracle as O, fixtures
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = fixtures.synthetic_problem(T=60, M=16, D=2, S=5)
    lo, hi = distributed.shard_range(5, rank, world)
    mine = copy.copy(prob)
    mine.X = prob.X[lo:hi]
    res = O.nll_and_grads(mine, collapsed=False)
    out = {k: torch.as_tensor(np.asarray(v)).clone() for k, v in res.items()}
    distributed.allreduce_shared(out)
    np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), lo=lo, hi=hi, **{k: v.numpy() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_process(tmp_path):
    from oracle import ffvd_oracle as O, fixtures
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    full = O.nll_and_grads(fixtures.synthetic_problem(T=60, M=16, D=2, S=5), collapsed=False)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for k in ("g_Z", "g_U", "g_logv", "g_logl", "g_logQ", "g_C", "g_d", "g_logR"):
        for p in parts:                            # every rank holds the full sum after the all-reduce
            assert np.max(np.abs(p[k] - full[k])) <= 1e-12 * max(1.0, np.max(np.abs(full[k]))), k
    gX = np.concatenate([p["g_X"] for p in parts])   # per-sample outputs stay with their owner
    nll = np.concatenate([p["nll"] for p in parts])
    assert np.max(np.abs(gX - full["g_X"])) <= 1e-13
    assert np.max(np.abs(nll - full["nll"])) <= 1e-13
