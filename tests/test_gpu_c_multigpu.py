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
