"""CPU: host-side logic of the sharded search -- page partition, global ids, candidate exchange and
merge order -- on a world_size-2 gloo group, with the oracle's top-k / merge injected in place of the
CUDA kernels (the product default needs a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_partition_pages(evdr):
    assert evdr.partition_pages(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert evdr.partition_pages(1_000_000, 8)[-1] == (875_000, 1_000_000)
    for n, w in [(7, 8), (0, 2), (13, 1)]:
        parts = evdr.partition_pages(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def _cpu_topk(scores, k, id_offset):
    n = scores.shape[1]
    order = sorted
    vals = torch.full((scores.shape[0], k), float("-inf"))
    ids = torch.full((scores.shape[0], k), -1, dtype=torch.int64)
    for q in range(scores.shape[0]):
        o = sorted(range(n), key=lambda c: (-scores[q, c].item(), c))[:k]
        vals[q, :len(o)] = scores[q, o]
        ids[q, :len(o)] = torch.tensor(o) + id_offset
    return vals, ids


def _cpu_merge(cs, ci, k):
    vals = torch.full((cs.shape[0], k), float("-inf"))
    ids = torch.full((cs.shape[0], k), -1, dtype=torch.int64)
    for q in range(cs.shape[0]):
        cand = sorted((( -cs[q, j].item(), ci[q, j].item()) for j in range(cs.shape[1]) if ci[q, j] >= 0))[:k]
        for t, (ns, i) in enumerate(cand):
            vals[q, t], ids[q, t] = -ns, i
    return vals, ids


def _worker(rank, world, port, n_pages, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import evdr_b200 as evdr
        g = torch.Generator().manual_seed(123)
        scores = torch.randn(5, n_pages, generator=g)
        scores[:, 3] = scores[:, n_pages - 2]          # a tie across shards
        lo, hi = evdr.partition_pages(n_pages, world)[rank]
        ss = evdr.ShardedSearch((lo, hi), mode="torchdist", local_topk=_cpu_topk, merge=_cpu_merge)
        vals, ids = ss.topk(scores[:, lo:hi].contiguous(), k)
        want_v, want_i = _cpu_topk(scores, k, 0)
        ok = torch.equal(ids, want_i) and torch.equal(vals, want_v)
        gathered = [None] * world
        dist.all_gather_object(gathered, ids.tolist())
        ok = ok and all(x == gathered[0] for x in gathered)      # identical on every rank
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pages,k", [(101, 10), (7, 5)])
def test_sharded_topk_world2_gloo(n_pages, k):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pages, k, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)


def _gather_worker(rank, world, port, n_pages, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import evdr_b200 as evdr
        g = torch.Generator().manual_seed(7)
        full = torch.randn(4, n_pages, generator=g)              # identical on both ranks
        w = torch.randn(4, n_pages, generator=g)
        lo, hi = evdr.partition_pages(n_pages, world)[rank]
        local = full[:, lo:hi].clone().requires_grad_(True)
        rows = evdr.gather_score_rows(local, n_pages, mode="torchdist")
        ok = torch.equal(rows, full)
        (rows * w).sum().backward()                               # d/d(local) = the local column slice of w
        ok = ok and torch.equal(local.grad, w[:, lo:hi])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pages", [10, 7, 1001])
def test_gather_score_rows_world2_gloo(n_pages):
    """Training collective: full (B, Np) rows on every rank from per-rank column blocks of
    unequal width; the backward is the local slice."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_pages, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)


def _check_worker(rank, world, port, ret):
    """bench.py's self-check at N > 1 (check_sharded_result) on a gloo group: accepts the correct merged top-k,
    rejects wrong ids, wrong scores and a rank that disagrees."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        import evdr_b200 as evdr
        npg, k = 50, 10
        scores = torch.randn(9, npg * world, generator=torch.Generator().manual_seed(5))
        lo, hi = rank * npg, (rank + 1) * npg
        local = scores[:, lo:hi].contiguous()
        ss = evdr.ShardedSearch((lo, hi), mode="torchdist", local_topk=_cpu_topk, merge=_cpu_merge)
        vals, ids = ss.topk(local, k)
        dev = torch.device("cpu")
        ok = bench.check_sharded_result(torch, dist, (vals, ids), local, npg, rank, world, dev) == "ok"
        bad_ids = ids.clone()
        bad_ids[0, 0] = (bad_ids[0, 0] + 1) % (npg * world)
        ok = ok and bench.check_sharded_result(torch, dist, (vals, bad_ids), local, npg, rank, world, dev) != "ok"
        ok = ok and bench.check_sharded_result(torch, dist, (vals + 1.0, ids), local, npg, rank, world, dev) != "ok"
        if rank == 1:                                   # only one rank holds a different (self-consistent) answer
            swapped = ids.clone()
            swapped[-1, [0, 1]] = swapped[-1, [1, 0]]
            sv = vals.clone()
            sv[-1, [0, 1]] = sv[-1, [1, 0]]
            res = bench.check_sharded_result(torch, dist, (sv, swapped), local, npg, rank, world, dev)
        else:
            res = bench.check_sharded_result(torch, dist, (vals, ids), local, npg, rank, world, dev)
        ok = ok and res != "ok"                         # rank 1: its own answer is wrong; rank 0: the ranks disagree
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_bench_sharded_check_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_check_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)
