"""GPU, multi-process: page-sharded search with the NCCL all-gather merge inside libevdr.so
(evdr_comm_init / evdr_sharded_topk), one process per GPU.  Needs >= 2 GPUs; skipped otherwise."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import evdr_b200 as evdr
        evdr.init_nccl_comm(dev)
        n_pages, nq, k = 1003, 37, 10
        # identical corpus + queries on every rank (CPU generator), each rank keeps its page range
        P, pm = evdr.synth.make_pages(n_pages, 200, seed=5, n_holes=3, empty_frac=0.02, min_len=30)
        Q, qm, rel = evdr.synth.make_queries(P, pm, nq, 32, seed=6, sigma=2.0, min_len=8)
        lo, hi = evdr.partition_pages(n_pages, world)[rank]
        shard = evdr.PackedPages.from_padded(P[lo:hi].to(dev), pm[lo:hi].to(dev), varlen=True, page_offset=lo)
        local = shard.score(Q.to(dev), qm.to(dev))
        top_s, top_i = evdr.ShardedSearch((lo, hi), mode="nccl").topk(local, k)
        # single-GPU answer over the whole corpus
        full = evdr.PackedPages.from_padded(P.to(dev), pm.to(dev), varlen=True).score(Q.to(dev), qm.to(dev))
        ref_s, ref_i = evdr.top_k(full, k)
        ok = torch.equal(top_i, ref_i) and torch.allclose(top_s, ref_s, rtol=0, atol=1e-5)
        ok = ok and bool((top_i[:, 0].cpu() == rel).float().mean() > 0.9)
        # the same search through the peer-memory exchange (one kernel, no NCCL on the data path):
        # repeated calls alternate the two epochs of the exchange regions, query counts above the grid
        # size exercise the CTA-strided loop
        evdr.init_peer_exchange(dev, max_candidates=1200 * 10)
        peer = evdr.ShardedSearch((lo, hi), mode="peer")
        for _ in range(5):
            ps, pi = peer.topk(local, k)
            ok = ok and torch.equal(pi, top_i) and torch.equal(ps, top_s)
        big = local.repeat(30, 1)[:1100]                      # 1100 query rows
        ns, ni = evdr.ShardedSearch((lo, hi), mode="nccl").topk(big, k)
        ps, pi = peer.topk(big, k)
        ok = ok and torch.equal(pi, ni) and torch.equal(ps, ns)
        ps, pi = peer.topk(local[:1], 3)
        ok = ok and torch.equal(pi, top_i[:1, :3])
        # fused search (scan with the top-k in its epilogue + exchange in one library call) on uniform
        # 1040-slot pages: 1, 3 and 7 queries take the fused route, 21 the scores + top-k route; both
        # exchange modes; bit-identical to the single-GPU top-k of the full score rows; graph replay too
        toks = evdr.synth.make_uniform_corpus(2001, 1040, device=dev, seed=77)          # same on every rank
        pmu = torch.ones(2001, 1040, dtype=torch.bool, device=dev)
        ulo, uhi = evdr.partition_pages(2001, world)[rank]
        ushard = evdr.PackedPages.from_uniform_tokens(toks[ulo:uhi].contiguous(), page_offset=ulo)
        whole = evdr.PackedPages.from_uniform_tokens(toks)
        for nq_f in (1, 3, 7, 21):
            Qf, qmf, _ = evdr.synth.make_queries(toks, pmu, nq_f, 32, seed=90 + nq_f, sigma=2.0, min_len=8, drop=0.1)
            ref_s, ref_i = evdr.top_k(whole.score(Qf, qmf), k)
            for mode in ("nccl", "peer"):
                ss = evdr.ShardedSearch((ulo, uhi), mode=mode)
                fs, fi = ss.search(ushard, Qf, qmf, k)
                ok = ok and torch.equal(fi, ref_i) and torch.equal(fs, ref_s)
                if nq_f in (3, 21):                             # k > 32: the unfused route through both exchanges
                    r40_s, r40_i = evdr.top_k(whole.score(Qf, qmf), 40)
                    f40_s, f40_i = ss.search(ushard, Qf, qmf, 40)
                    ok = ok and torch.equal(f40_i, r40_i) and torch.equal(f40_s, r40_s)
                if nq_f == 1:
                    gs = evdr.GraphedSearch(ss, ushard, nq_f, k)
                    for _ in range(3):
                        g_s, g_i = gs(Qf.float(), qmf)
                        ok = ok and torch.equal(g_i, ref_i) and torch.equal(g_s, ref_s)
        # training collective: full (nq, n_pages) rows from the per-rank column blocks (unequal widths),
        # gradient of the local block = the local column slice
        loc = local.clone().requires_grad_(True)
        rows = evdr.gather_score_rows(loc, n_pages)
        ok = ok and torch.allclose(rows, full, rtol=0, atol=1e-5)
        wgt = torch.randn(nq, n_pages, generator=torch.Generator().manual_seed(1)).to(dev)
        (rows * wgt).sum().backward()
        ok = ok and torch.equal(loc.grad, wgt[:, lo:hi])
        gathered = [None] * world
        dist.all_gather_object(gathered, top_i.cpu().tolist())
        ok = ok and all(g == gathered[0] for g in gathered)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_search_nccl():
    world = min(torch.cuda.device_count(), 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _timeout_worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), EVDR_PEER_TIMEOUT_MS="300")
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import evdr_b200 as evdr
        evdr.init_nccl_comm(dev)
        evdr.init_peer_exchange(dev, max_candidates=64 * 16)
        lo, hi = evdr.partition_pages(400, world)[rank]
        s = torch.randn(3, hi - lo, generator=torch.Generator().manual_seed(rank)).to(dev)
        peer = evdr.ShardedSearch((lo, hi), mode="peer")
        ok = True
        if rank == 0:
            # rank 1 never makes this call: the wait is bounded, the kernel completes with padding instead of
            # trapping, and the NEXT peer call on this handle reports the failure
            top_s, top_i = peer.topk(s, 5)
            torch.cuda.synchronize(dev)
            ok = ok and bool((top_i == -1).all()) and bool(torch.isinf(top_s).all())
            try:
                peer.topk(s, 5)
                ok = False
            except evdr._lib.EvdrError as e:
                ok = ok and "timed out" in str(e)
        dist.barrier()
        # the process and the context are alive: the NCCL exchange still works on both ranks
        ns, ni = evdr.ShardedSearch((lo, hi), mode="nccl").topk(s, 5)
        full = [torch.empty_like(s) for _ in range(world)]
        dist.all_gather(full, s)
        ws, wi = evdr.top_k(torch.cat(full, dim=1), 5)
        ok = ok and torch.equal(ni, wi) and torch.equal(ns, ws)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_exchange_timeout_is_reported_not_trapped():
    """a rank that never arrives must not kill the others' CUDA context."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_timeout_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(200)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1), dict(ret)


def test_peer_exchange_single_rank(evdr, cuda):
    """world = 1: the exchange kernel pushes into its own region, waits on its own flags and merges --
    the result must equal evdr_topk, including (-inf, -1) padding when k exceeds the page count."""
    import ctypes as C
    lib = evdr._lib.load()
    h = evdr._lib.handle(cuda)
    if lib.evdr_peer_world(h) == 0:
        buf = (C.c_ubyte * 64)()
        evdr._lib.check(lib.evdr_peer_region_init(h, 0, 1, 300 * 16, C.cast(buf, C.c_void_p)), "init")
        evdr._lib.check(lib.evdr_peer_connect(h, C.cast(buf, C.c_void_p)), "connect")
    assert lib.evdr_peer_world(h) == 1
    g = torch.Generator(device=cuda).manual_seed(3)
    for nq, npg, k in ((7, 500, 10), (300, 37, 16), (1, 5, 8)):
        s = torch.randn(nq, npg, generator=g, device=cuda)
        s[:, 1] = s[:, 0]                                      # ties resolve by id ascending
        want_s, want_i = evdr.top_k(s, k, id_offset=100)
        got_s, got_i = evdr.ShardedSearch((100, 100 + npg), mode="peer").topk(s, k)
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
    with pytest.raises(ValueError):
        evdr.ShardedSearch((0, 10), mode="peer").topk(torch.zeros(400, 10, device=cuda), 16)   # > region


def test_gather_score_rows_single_rank(evdr, cuda):
    s = torch.randn(6, 45, device=cuda)
    wide = torch.zeros(6, 64, device=cuda)
    wide[:, :45] = s
    loc = wide[:, :45].requires_grad_(True)                   # strided rows
    rows = evdr.gather_score_rows(loc, 45)
    assert torch.equal(rows, s)
    rows.sum().backward()
    assert torch.equal(loc.grad, torch.ones_like(s))
