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


def _worker_time(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import copy
    from ffvd_b200 import distributed, _capi
    from oracle import ffvd_oracle as O, fixtures
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = fixtures.synthetic_problem(T=61, M=16, D=2, S=1)
    pd = {k: getattr(prob, k) for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")}

    def evaluate(blk, out, extra):           # the oracle stands in for ctx.nll_grads on a block
        q = copy.copy(prob)
        q.X, q.Y, q.ctrl = blk["X"], blk["Y"], blk["ctrl"]
        res = O.nll_and_grads(q, collapsed=False, shared_priors=not (extra & _capi.FLAG_NO_SHARED_PRIORS),
                              x0_prior=not (extra & _capi.FLAG_NO_X0_PRIOR))
        for k in list(out):
            out[k] = torch.as_tensor(np.atleast_1d(np.asarray(res[k]))).clone().reshape(out[k].shape if out[k] is not None else -1)

    blk, a, b = distributed.time_block(pd, rank, world)
    out = {"nll": torch.zeros(1, dtype=torch.float64), "terms": torch.zeros(6, dtype=torch.float64),
           "g_X": torch.zeros(b - a + 1, 2, dtype=torch.float64)}
    for k in ("Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
        out["g_" + k] = torch.zeros(np.asarray(pd[k]).shape, dtype=torch.float64)
    distributed.evaluate_time_sharded(evaluate, pd, out, rank, world)
    Xb = torch.as_tensor(np.array(blk["X"]))          # halo refresh after an owner-side update of row 0
    Xb[0] += 1.0 + rank
    distributed.exchange_halo_row(Xb, rank, world)
    np.savez(os.path.join(tmpdir, "trank%d.npz" % rank), a=a, b=b, halo=Xb[-1].numpy(), **{k: v.numpy() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_time_sharded_equals_single_process(tmp_path):
    """One trajectory, T split over two ranks with a one-row halo: nll, terms, shared gradients and the stitched x-bar equal
    the single-process evaluation (the S < #GPUs case)."""
    from oracle import ffvd_oracle as O, fixtures
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_time, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    prob = fixtures.synthetic_problem(T=61, M=16, D=2, S=1)
    full = O.nll_and_grads(prob, collapsed=False)
    parts = [np.load(os.path.join(str(tmp_path), "trank%d.npz" % r)) for r in range(world)]
    for p in parts:
        assert abs(p["nll"][0] - full["nll"]) <= 1e-13 * abs(full["nll"])
        assert np.max(np.abs(p["terms"] - full["terms"])) <= 1e-13
        for k in ("g_Z", "g_U", "g_logv", "g_logl", "g_logQ", "g_C", "g_d", "g_logR"):
            assert np.max(np.abs(p[k] - full[k])) <= 1e-12 * max(1.0, np.max(np.abs(full[k]))), k
    # owners: rank r holds final rows a..b-1 (the last rank also row b)
    gX = np.concatenate([parts[0]["g_X"][:-1], parts[1]["g_X"]])
    assert gX.shape == full["g_X"].shape
    assert np.max(np.abs(gX - full["g_X"])) <= 1e-13 * max(1.0, np.max(np.abs(full["g_X"])))
    # halo refresh: rank 0's halo row is rank 1's updated first row
    assert np.allclose(parts[0]["halo"], prob.X[int(parts[1]["a"])] + 2.0)
