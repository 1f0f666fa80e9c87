"""Multi-GPU retrieval: the page corpus is sharded by contiguous page range across the ranks of
one box (one process per GPU), queries are replicated, every rank computes a local top-k and the
candidates are merged after ONE all-gather.  There is no other data-path collective.

Three exchange modes:
* "peer": `evdr_peer_sharded_topk` -- local top-k, then ONE kernel that pushes the candidates into
  every rank's exchange region with remote stores over NVLink (CUDA IPC peer memory), signals per-query
  flags and merges as soon as all ranks' candidates of a query have landed; no NCCL on the data path
  (set up with `init_peer_exchange`);
* "nccl"  (default on CUDA): `evdr_sharded_topk` -- local top-k, ncclAllGather of nq*k (score, id)
  pairs on the caller's stream and the merge kernel, all inside libevdr.so;
* "torchdist": the same three steps with the all-gather done by torch.distributed on whatever
  backend the process group has.  This is what the world_size-2 gloo tests exercise on CPU with the
  oracle's top-k / merge injected in place of the CUDA kernels (host logic only).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def partition_pages(n_pages: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous page ranges [start, end) per rank; the first n_pages % world ranks get one extra."""
    base, rem = divmod(n_pages, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def init_nccl_comm(device: torch.device, group=None) -> None:
    """Create libevdr's NCCL communicator for this rank: rank 0 makes the unique id, torch.distributed
    (the plumbing) broadcasts it, every rank calls evdr_comm_init."""
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        _lib.check(lib.evdr_comm_unique_id(C.cast(buf, C.c_void_p)), "evdr_comm_unique_id")
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    backend = dist.get_backend(group)
    t = uid.to(device) if backend == "nccl" else uid
    dist.broadcast(t, src=0, group=group)
    raw = bytes(t.cpu().tolist())
    buf = (C.c_ubyte * 128).from_buffer_copy(raw)
    _lib.check(lib.evdr_comm_init(_lib.handle(device), C.cast(buf, C.c_void_p), rank, world), "evdr_comm_init")


def init_peer_exchange(device: torch.device, group=None, max_candidates: int = 64 * 128) -> None:
    """Set up libevdr's peer-memory exchange for this rank: allocate the exchange region, gather every
    rank's CUDA IPC handle with torch.distributed (the plumbing) and open the peers' regions.  After this
    ShardedSearch(mode="peer") needs no NCCL call on the data path.  `max_candidates` bounds nq * k per
    search (default: 64 queries x top-128, or 819 queries x top-10).

    Collective and failure-consistent: every rank takes part in both exchanges even when its own step
    failed, and either all ranks return or all raise (so callers can fall back to mode="nccl" together)."""
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    h = _lib.handle(device)
    if lib.evdr_peer_world(h) == world:
        return
    on_dev = dist.get_backend(group) == "nccl"

    def gather_bytes(payload: bytes):
        mine = torch.tensor(list(payload), dtype=torch.uint8)
        mine = mine.to(device) if on_dev else mine
        allt = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine, group=group)
        return [bytes(t.cpu().tolist()) for t in allt]

    err = None
    buf = (C.c_ubyte * 64)()
    try:
        _lib.check(lib.evdr_peer_region_init(h, rank, world, int(max_candidates), C.cast(buf, C.c_void_p)),
                   "evdr_peer_region_init")
    except Exception as e:          # still take part in the exchange below
        err = e
    got = gather_bytes(bytes(buf) + (b"\x01" if err is None else b"\x00"))
    if err is None and all(g[64] == 1 for g in got):
        raw = b"".join(g[:64] for g in got)
        blob = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        try:
            _lib.check(lib.evdr_peer_connect(h, C.cast(blob, C.c_void_p)), "evdr_peer_connect")
        except Exception as e:
            err = e
    elif err is None:
        err = _lib.EvdrError("a peer rank could not allocate its exchange region")
    ok = gather_bytes(b"\x01" if err is None else b"\x00")   # also the barrier: nobody pushes into a region
    if not all(o[0] == 1 for o in ok):                          # before its owner has opened everyone
        raise _lib.EvdrError(f"peer-memory exchange unavailable: {err or 'a peer rank failed to connect'}")


class ShardedSearch:
    """Global top-k over per-rank score blocks.

    local_topk(scores, k, id_offset) -> (top_scores, top_ids) and merge(cand_scores, cand_ids, k) ->
    (top_scores, top_ids) default to the CUDA kernels; tests inject CPU doubles.
    """

    def __init__(self, page_range: Tuple[int, int], *, mode: str = "nccl", group=None,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.start, self.end = page_range
        self.mode = mode
        self.group = group
        if local_topk is None or merge is None:
            from .topk import topk as _topk, topk_merge as _merge
            local_topk = local_topk or (lambda s, k, off: _topk(s, k, id_offset=off))
            merge = merge or _merge
        self._local_topk = local_topk
        self._merge = merge

    def topk(self, local_scores: torch.Tensor, k: int):
        """local_scores (nq, n_local_pages) fp32 -> global (scores (nq,k), ids (nq,k)) on every rank."""
        nq, n_local = local_scores.shape
        assert n_local == self.end - self.start, "score block does not match this rank's page range"
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if self.mode in ("nccl", "peer"):
            dev = local_scores.device
            _lib.require_cuda(local_scores, "local_scores")
            top_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
            top_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
            if local_scores.stride(1) != 1:
                local_scores = local_scores.contiguous()
            lib = _lib.load()
            fn = lib.evdr_peer_sharded_topk if self.mode == "peer" else lib.evdr_sharded_topk
            _lib.check(fn(_lib.handle(dev), local_scores.data_ptr(), local_scores.stride(0), nq, n_local, k,
                          self.start, top_s.data_ptr(), top_i.data_ptr(), _lib.stream_ptr(dev)),
                       "evdr_peer_sharded_topk" if self.mode == "peer" else "evdr_sharded_topk")
            return top_s, top_i
        loc_s, loc_i = self._local_topk(local_scores, k, self.start)
        if world == 1:
            return loc_s, loc_i
        gs = [torch.empty_like(loc_s) for _ in range(world)]
        gi = [torch.empty_like(loc_i) for _ in range(world)]
        dist.all_gather(gs, loc_s.contiguous(), group=self.group)
        dist.all_gather(gi, loc_i.contiguous(), group=self.group)
        return self._merge(torch.cat(gs, dim=1), torch.cat(gi, dim=1), k)


    # ------------------------------------------------------------------ fused search
    _EXCHANGE = {"none": 0, "nccl": 1, "peer": 2}

    def search_packed(self, pk, Qb: torch.Tensor, qm32: torch.Tensor, k: int,
                      out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """Packed queries against this rank's shard `pk` -> global top-k on every rank in ONE library call
        (evdr_sharded_search): local scan with the top-k folded into its epilogue (<= 8 queries on long pages:
        only k candidates per CTA leave the SM, no (nq, np) score row), candidate exchange, merge."""
        if pk.k_blocks != 1:
            raise NotImplementedError("fused search runs on plain bf16 shards")
        if pk.n_pages != self.end - self.start:
            raise ValueError("shard does not match this rank's page range")
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        mode = self.mode if world > 1 or self.mode == "peer" else "none"
        if mode not in self._EXCHANGE:
            raise ValueError(f"search_packed needs mode 'nccl' or 'peer', not {self.mode!r}")
        nq = Qb.shape[0]
        dev = pk.device
        top_s, top_i = out if out is not None else (torch.empty((nq, k), dtype=torch.float32, device=dev),
                                                     torch.empty((nq, k), dtype=torch.int64, device=dev))
        lib = _lib.load()
        _lib.check(lib.evdr_sharded_search(_lib.handle(dev), Qb.data_ptr(), qm32.data_ptr(), nq, pk.tokens.data_ptr(),
                                           pk.tokens.shape[0], pk.chunk_off.data_ptr(), pk.n_pages, pk.total_chunks,
                                           k, self.start, self._EXCHANGE[mode], top_s.data_ptr(), top_i.data_ptr(),
                                           _lib.stream_ptr(dev)), "evdr_sharded_search")
        return top_s, top_i

    def search(self, pk, Q: torch.Tensor, qmask: torch.Tensor, k: int):
        """(nq, lq <= 32, 128) queries -> global top-k (scores (nq, k), ids (nq, k) int64) on every rank."""
        from .corpus import pack_queries
        if Q.shape[1] > _lib.TC_LQ:
            return self.topk(pk.score(Q, qmask), k)            # long queries: 32-token groups, unfused
        Qb, qm32 = pack_queries(Q, qmask)
        return self.search_packed(pk, Qb, qm32, k)


class GraphedSearch:
    """pack -> scan (+ fused top-k) -> exchange -> merge captured ONCE as a CUDA graph and replayed per query
    batch: the latency path of a sharded corpus (1 M pages over 8 GPUs: a 4.5 ms scan, where four launches
    and their host-side issue would otherwise show).  Inputs are copied into static buffers, results are
    returned in static buffers (valid until the next call).  Collective at world > 1: every rank replays."""

    def __init__(self, search: "ShardedSearch", pk, nq: int, k: int = 10, lq: int = _lib.TC_LQ):
        from .corpus import pack_queries
        dev = pk.device
        self.search, self.pk, self.k = search, pk, k
        self.Q = torch.zeros((nq, lq, _lib.TC_DIM), dtype=torch.float32, device=dev)
        self.qmask = torch.zeros((nq, lq), dtype=torch.bool, device=dev)
        self.top_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        self.top_i = torch.empty((nq, k), dtype=torch.int64, device=dev)

        def body():
            Qb, qm32 = pack_queries(self.Q, self.qmask)
            search.search_packed(pk, Qb, qm32, k, out=(self.top_s, self.top_i))

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                    # warm-up outside capture (lazy module loads, pool growth)
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()

    def __call__(self, Q: torch.Tensor, qmask: torch.Tensor):
        """Q (nq, lq <= the captured lq, 128) on the host (pinned for an asynchronous copy) or the device."""
        nq, lq = Q.shape[0], Q.shape[1]
        if nq != self.Q.shape[0] or lq > self.Q.shape[1] or Q.shape[2] != self.Q.shape[2]:
            raise ValueError(f"graph captured for queries {tuple(self.Q.shape)}, got {tuple(Q.shape)}")
        if lq == self.Q.shape[1]:
            self.Q.copy_(Q, non_blocking=True)
            self.qmask.copy_(qmask != 0 if qmask.dtype != torch.bool else qmask, non_blocking=True)
        else:                                   # shorter queries: the remaining token slots are masked out
            self.qmask.zero_()
            self.Q[:, :lq].copy_(Q, non_blocking=True)
            self.qmask[:, :lq].copy_(qmask != 0 if qmask.dtype != torch.bool else qmask, non_blocking=True)
        self.graph.replay()
        return self.top_s, self.top_i


class _GatherScoreRows(torch.autograd.Function):
    """(B, n_local) score block of this rank's pages -> full (B, n_total) rows on every rank; the backward
    is the local column slice of the incoming gradient (each rank owns its pages' parameters: no reduce)."""

    @staticmethod
    def forward(ctx, local, n_total, mode, group):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        start, end = partition_pages(n_total, world)[rank]
        if local.shape[1] != end - start:
            raise ValueError(f"rank {rank} owns pages [{start}, {end}) but the score block has {local.shape[1]} columns")
        ctx.range = (start, end)
        B = local.shape[0]
        src = local.detach().to(torch.float32)
        if mode == "nccl":
            _lib.require_cuda(src, "local_scores")
            if src.stride(1) != 1:
                src = src.contiguous()
            out = torch.empty((B, n_total), dtype=torch.float32, device=src.device)
            lib = _lib.load()
            _lib.check(lib.evdr_allgather_score_rows(_lib.handle(src.device), src.data_ptr(), src.stride(0), B, n_total,
                                                     out.data_ptr(), out.stride(0), _lib.stream_ptr(src.device)),
                       "evdr_allgather_score_rows")
            return out
        # torch.distributed on whatever backend the group has (gloo in the CPU tests): equal-size blocks
        cmax = partition_pages(n_total, world)[0][1]
        pad = torch.zeros((B, cmax), dtype=torch.float32, device=src.device)
        pad[:, : end - start] = src
        blocks = [torch.empty_like(pad) for _ in range(world)]
        if world > 1:
            dist.all_gather(blocks, pad, group=group)
        else:
            blocks = [pad]
        return torch.cat([blk[:, : e - s] for blk, (s, e) in zip(blocks, partition_pages(n_total, world))], dim=1)

    @staticmethod
    def backward(ctx, g):
        s, e = ctx.range
        return g[:, s:e].contiguous(), None, None, None


def gather_score_rows(local_scores: torch.Tensor, n_total: int, *, mode: str = "nccl", group=None) -> torch.Tensor:
    """Training collective of SURVEY.md 8e: full (B, n_total) score rows on every rank from per-rank column
    blocks (pages partitioned by `partition_pages`), differentiable -- the gradient of a rank's block is the
    matching column slice.  mode="nccl": ncclAllGather inside libevdr.so; "torchdist": torch.distributed."""
    return _GatherScoreRows.apply(local_scores, int(n_total), mode, group)
