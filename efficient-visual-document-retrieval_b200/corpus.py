"""Packed page corpus: the HBM layout of the tensor-core MaxSim path.

A corpus is a flat bf16 token stream ``tokens[n_tokens, 128]`` in which page ``c`` owns the
16-token chunks ``chunk_off[c] .. chunk_off[c+1]``.  Masked / padded slots inside a page hold a
duplicate of one of the page's valid tokens, so the scorer needs no mask at all; a page without
valid tokens is all-zero tokens and scores exactly 0 like evaluator/retrieval.py:192,204.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import CHUNK, TC_DIM, TC_LQ, TC_LQ_MAX, TOPK_MAX


def _mask_u8(mask: torch.Tensor) -> torch.Tensor:
    """Masks of any dtype -> contiguous uint8, non-zero = valid (reference `.bool()`, :179-180)."""
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    return (mask != 0).contiguous().view(torch.uint8)


def _float_operand(t: torch.Tensor) -> torch.Tensor:
    """fp32 / bf16 are consumed natively; other float dtypes are up-cast like the reference (:176-177)."""
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.to(torch.float32)
    return t.contiguous()


def split_long_queries(Q: torch.Tensor, qmask: torch.Tensor):
    """Queries longer than 32 tokens -> G = ceil(lq/32) groups of 32 token slots each, returned as
    (Q (nq*G, 32, d), qmask (nq*G, 32), G).  Exact for MaxSim: the max over page tokens is taken per
    query token and the per-group scores are added back with evdr_group_sum."""
    nq, lq, d = Q.shape
    G = (lq + TC_LQ - 1) // TC_LQ
    if G == 1:
        return Q, qmask, 1
    Qp = torch.zeros((nq, G * TC_LQ, d), dtype=Q.dtype, device=Q.device)
    mp = torch.zeros((nq, G * TC_LQ), dtype=torch.bool, device=Q.device)
    Qp[:, :lq] = Q
    mp[:, :lq] = qmask != 0
    return Qp.view(nq * G, TC_LQ, d), mp.view(nq * G, TC_LQ), G


def group_sum(scores: torch.Tensor, groups: int) -> torch.Tensor:
    """(nq*groups, np) -> (nq, np): out[q] = sum_g scores[q*groups + g] (evdr_group_sum)."""
    if groups == 1:
        return scores
    rows, npg = scores.shape
    nq = rows // groups
    out = torch.empty((nq, npg), dtype=torch.float32, device=scores.device)
    lib = _lib.load()
    _lib.check(lib.evdr_group_sum(_lib.handle(scores.device), scores.data_ptr(), scores.stride(0), nq, groups,
                                  npg, out.data_ptr(), out.stride(0), _lib.stream_ptr(scores.device)),
               "evdr_group_sum")
    return out


def pack_queries(Q: torch.Tensor, qmask: torch.Tensor, k_blocks: int = 1):
    """(nq, lq<=32, 128) -> bf16 (nq, 32, 128) with masked slots zeroed, uint8 (nq, 32).
    k_blocks=3 writes the [hi | lo] rows (256 wide) of the bf16x3 split."""
    _lib.require_cuda(Q, "Q")
    nq, lq, d = Q.shape
    Qc = _float_operand(Q)
    qm = _mask_u8(qmask)
    dev = Q.device
    Qb = torch.empty((nq, TC_LQ, TC_DIM * _lib.row_blocks(k_blocks)), dtype=torch.bfloat16, device=dev)
    qm32 = torch.empty((nq, TC_LQ), dtype=torch.uint8, device=dev)
    lib = _lib.load()
    _lib.check(lib.evdr_pack_queries(_lib.handle(dev), Qc.data_ptr(), _lib.dtype_code(Qc), qm.data_ptr(),
                                     nq, lq, d, k_blocks, Qb.data_ptr(), qm32.data_ptr(), _lib.stream_ptr(dev)),
               "evdr_pack_queries")
    return Qb, qm32


class PackedPages:
    """Device-resident packed corpus shard."""

    def __init__(self, tokens: torch.Tensor, chunk_off: torch.Tensor, n_pages: int, total_chunks: int,
                 page_offset: int = 0, k_blocks: int = 1):
        self.k_blocks = int(k_blocks)   # 1 = bf16 rows, 3 = bf16x3 split: rows [hi | lo]
        self.tokens = tokens            # bf16 [n_tokens, 128] or, for the split, [n_tokens, 256]
        self.chunk_off = chunk_off      # int32 [n_pages + 1]
        self.n_pages = int(n_pages)
        self.total_chunks = int(total_chunks)
        self.page_offset = int(page_offset)   # global id of local page 0 (sharded corpora)
        self.device = tokens.device

    # ------------------------------------------------------------------ builders
    @staticmethod
    def from_padded(P: torch.Tensor, pmask: torch.Tensor, *, normalize: bool = False,
                    varlen: bool = False, page_offset: int = 0, k_blocks: int = 1) -> "PackedPages":
        """Pack the reference's padded layout: P (np, lp, 128) fp32|bf16, pmask (np, lp) any dtype.

        varlen=False sizes every page at ceil(lp/16) chunks (no host sync; right for the per-step
        student tensor).  varlen=True trims each page after its last valid token (one device->host
        read of the total; right for a corpus that is packed once).
        normalize=True fuses l2_normalize(P * pmask[..., None]) (mainv2_iter_listwise.py:280).
        k_blocks=3 stores the bf16x3 split (2x the bytes, 3x the tensor work, near-fp32 similarities).
        """
        _lib.require_cuda(P, "P")
        npg, lp, d = P.shape
        if d != TC_DIM:
            raise NotImplementedError(f"tensor-core path needs d=128, got {d}")
        Pc = _float_operand(P)
        pm = _mask_u8(pmask)
        dev = P.device
        lib = _lib.load()
        h = _lib.handle(dev)
        st = _lib.stream_ptr(dev)
        chunk_off = torch.empty(npg + 1, dtype=torch.int32, device=dev)
        _lib.check(lib.evdr_pack_plan(h, pm.data_ptr(), npg, lp, 1 if varlen else 0, chunk_off.data_ptr(), st),
                   "evdr_pack_plan")
        if varlen:
            total = int(chunk_off[-1].item())
        else:
            total = npg * ((lp + CHUNK - 1) // CHUNK)
        tokens = torch.empty((total * CHUNK, TC_DIM * _lib.row_blocks(k_blocks)), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.evdr_pack_pages(h, Pc.data_ptr(), _lib.dtype_code(Pc), pm.data_ptr(),
                                       chunk_off.data_ptr(), npg, lp, d, 1 if normalize else 0, k_blocks,
                                       tokens.data_ptr(), st), "evdr_pack_pages")
        return PackedPages(tokens, chunk_off, npg, total, page_offset, k_blocks)

    @staticmethod
    def from_object_arrays(documents, doc_attnmask=None, doc_imgmask=None, *, device, normalize: bool = False,
                           page_offset: int = 0, k_blocks: int = 1) -> "PackedPages":
        """Pack the reference's on-disk layout directly: ``documents`` is a sequence (numpy object array /
        list) of (Li, 128) float arrays, the masks sequences of (Li,) bool-like arrays or None -- the
        arguments of utils/preprocess_data.py:80-105 (preprocess_docs).  The page mask is
        valid & attn & img like :101.  Tokens are concatenated once on the host, copied once, and packed
        on the device (evdr_pack_*_ragged); no padded (N, Lmax, D) fp32 tensor is built."""
        import numpy as np
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.EvdrError("from_object_arrays packs on the GPU: this package has no CPU fallback")
        docs = [np.asarray(x, dtype=np.float32).reshape(-1, TC_DIM) for x in documents]
        npg = len(docs)
        lens = np.array([x.shape[0] for x in docs], dtype=np.int64)
        off = np.zeros(npg + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        total_tok = int(off[-1])
        flat = np.concatenate(docs, axis=0) if total_tok else np.zeros((0, TC_DIM), np.float32)
        mask = np.ones(total_tok, dtype=bool)
        for src in (doc_attnmask, doc_imgmask):
            if src is None:
                continue
            for i in range(npg):
                mi = src[i]
                if mi is None:
                    continue
                mi = np.asarray(mi).astype(bool).reshape(-1)
                L = min(int(lens[i]), mi.shape[0])
                seg = mask[off[i]:off[i + 1]]
                seg[:L] &= mi[:L]
                seg[L:] = False                    # mask shorter than the page: padded False (:75-76)
        tok_d = torch.from_numpy(flat).to(device, non_blocking=True)
        mask_d = torch.from_numpy(mask.view(np.uint8)).to(device, non_blocking=True)
        off_d = torch.from_numpy(off).to(device, non_blocking=True)
        if total_tok == 0:
            tok_d = torch.zeros((1, TC_DIM), dtype=torch.float32, device=device)
            mask_d = torch.zeros(1, dtype=torch.uint8, device=device)
        lib = _lib.load()
        h = _lib.handle(device)
        st = _lib.stream_ptr(device)
        chunk_off = torch.empty(npg + 1, dtype=torch.int32, device=device)
        _lib.check(lib.evdr_pack_plan_ragged(h, mask_d.data_ptr(), off_d.data_ptr(), npg, chunk_off.data_ptr(), st),
                   "evdr_pack_plan_ragged")
        total = int(chunk_off[-1].item())
        tokens = torch.empty((total * CHUNK, TC_DIM * _lib.row_blocks(k_blocks)), dtype=torch.bfloat16, device=device)
        _lib.check(lib.evdr_pack_pages_ragged(h, tok_d.data_ptr(), _lib.EVDR_F32, mask_d.data_ptr(), off_d.data_ptr(),
                                              chunk_off.data_ptr(), npg, max(1, int(lens.max()) if npg else 1),
                                              TC_DIM, 1 if normalize else 0, k_blocks, tokens.data_ptr(), st),
                   "evdr_pack_pages_ragged")
        return PackedPages(tokens, chunk_off, npg, total, page_offset, k_blocks)

    @staticmethod
    def from_tensor_list(pages, *, device, extra_zero_token=None, normalize: bool = False,
                         page_offset: int = 0, k_blocks: int = 1) -> "PackedPages":
        """Pack a list of (Li, 128) torch tensors (any float dtype, any device) without masks: one
        concatenation, at most one host->device copy, one ragged pack on the device -- no per-page copies and
        no padded (N, Lmax, D) intermediate.  extra_zero_token[i] appends one all-zero token to page i (how the
        list scorer's zero LEFT padding takes part in the max, evaluator/retrieval.py:30-45,136)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.EvdrError("from_tensor_list packs on the GPU: this package has no CPU fallback")
        npg = len(pages)
        lens = [int(p.shape[0]) for p in pages]
        if extra_zero_token is None:
            extra_zero_token = [False] * npg
        if len({p.device for p in pages}) > 1:
            pages = [p.to(device) for p in pages]
        parts, zero = [], None
        for p, z in zip(pages, extra_zero_token):
            parts.append(p.reshape(-1, TC_DIM))
            if z:
                if zero is None or zero.device != p.device or zero.dtype != p.dtype:
                    zero = torch.zeros((1, TC_DIM), dtype=p.dtype, device=p.device)
                parts.append(zero)
        flat = torch.cat(parts, dim=0).to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
        lens_t = torch.tensor([l + (1 if z else 0) for l, z in zip(lens, extra_zero_token)], dtype=torch.int64)
        off = torch.zeros(npg + 1, dtype=torch.int64)
        torch.cumsum(lens_t, 0, out=off[1:])
        total_tok = int(off[-1])
        max_len = int(lens_t.max()) if npg else 1
        off_d = off.to(device, non_blocking=True)
        if total_tok == 0:
            flat = torch.zeros((1, TC_DIM), dtype=torch.float32, device=device)
        mask_d = torch.ones(max(1, total_tok), dtype=torch.uint8, device=device)
        lib = _lib.load()
        h = _lib.handle(device)
        st = _lib.stream_ptr(device)
        chunk_off = torch.empty(npg + 1, dtype=torch.int32, device=device)
        _lib.check(lib.evdr_pack_plan_ragged(h, mask_d.data_ptr(), off_d.data_ptr(), npg, chunk_off.data_ptr(), st),
                   "evdr_pack_plan_ragged")
        # every page has at least one chunk and ceil(len/16) otherwise: known on the host, no device read
        total = sum(max(1, (int(l) + CHUNK - 1) // CHUNK) for l in lens_t.tolist())
        tokens = torch.empty((total * CHUNK, TC_DIM * _lib.row_blocks(k_blocks)), dtype=torch.bfloat16, device=device)
        _lib.check(lib.evdr_pack_pages_ragged(h, flat.data_ptr(), _lib.EVDR_F32, mask_d.data_ptr(), off_d.data_ptr(),
                                              chunk_off.data_ptr(), npg, max(1, max_len), TC_DIM,
                                              1 if normalize else 0, k_blocks, tokens.data_ptr(), st),
                   "evdr_pack_pages_ragged")
        return PackedPages(tokens, chunk_off, npg, total, page_offset, k_blocks)

    @staticmethod
    def from_uniform_tokens(tokens: torch.Tensor, page_offset: int = 0) -> "PackedPages":
        """Zero-copy view of an all-valid bf16 corpus (np, L, 128) with L % 16 == 0."""
        _lib.require_cuda(tokens, "tokens")
        npg, L, d = tokens.shape
        if tokens.dtype != torch.bfloat16 or d != TC_DIM or L % CHUNK != 0 or not tokens.is_contiguous():
            raise ValueError("from_uniform_tokens needs contiguous bf16 (np, L%16==0, 128)")
        cpp = L // CHUNK
        chunk_off = torch.arange(0, (npg + 1) * cpp, cpp, dtype=torch.int32, device=tokens.device)
        return PackedPages(tokens.view(npg * L, d), chunk_off, npg, npg * cpp, page_offset)

    # ------------------------------------------------------------------ scoring
    def score_packed(self, Qb: torch.Tensor, qm32: torch.Tensor, out: torch.Tensor | None = None,
                     pages_per_item: int = 0) -> torch.Tensor:
        """Tensor-core MaxSim of packed queries against this shard -> fp32 (nq, n_pages)."""
        nq = Qb.shape[0]
        dev = self.device
        if out is None:
            out = torch.empty((nq, self.n_pages), dtype=torch.float32, device=dev)
        lib = _lib.load()
        _lib.check(lib.evdr_maxsim_packed_fwd(
            _lib.handle(dev), Qb.data_ptr(), qm32.data_ptr(), nq, self.tokens.data_ptr(),
            self.tokens.shape[0], self.chunk_off.data_ptr(), self.n_pages, self.total_chunks,
            pages_per_item, self.k_blocks, out.data_ptr(), out.stride(0), _lib.stream_ptr(dev)),
            "evdr_maxsim_packed_fwd")
        return out

    def score(self, Q: torch.Tensor, qmask: torch.Tensor) -> torch.Tensor:
        """MaxSim of (nq, lq <= 128, 128) queries against this shard -> fp32 (nq, n_pages)."""
        Qs, ms, G = split_long_queries(Q, qmask)
        Qb, qm32 = pack_queries(Qs, ms, self.k_blocks)
        return group_sum(self.score_packed(Qb, qm32), G)

    def search(self, Q: torch.Tensor, qmask: torch.Tensor, k: int = 10):
        """Top-k pages per query: (scores (nq,k) fp32, ids (nq,k) int64 global page ids)."""
        from .topk import topk
        if self.k_blocks == 1 and Q.shape[1] <= TC_LQ and k <= TOPK_MAX:
            # one library call; for <= 8 queries on long pages the top-k is folded into the scan's epilogue
            Qb, qm32 = pack_queries(Q, qmask)
            nq = Qb.shape[0]
            top_s = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            top_i = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            lib = _lib.load()
            _lib.check(lib.evdr_sharded_search(_lib.handle(self.device), Qb.data_ptr(), qm32.data_ptr(), nq,
                                               self.tokens.data_ptr(), self.tokens.shape[0], self.chunk_off.data_ptr(),
                                               self.n_pages, self.total_chunks, k, self.page_offset, 0,
                                               top_s.data_ptr(), top_i.data_ptr(), _lib.stream_ptr(self.device)),
                       "evdr_sharded_search")
            return top_s, top_i
        scores = self.score(Q, qmask)
        return topk(scores, k, id_offset=self.page_offset)

    @property
    def nbytes(self) -> int:
        return self.tokens.numel() * 2 + self.chunk_off.numel() * 4


class HostCorpus:
    """A uniform bf16 corpus that lives in PINNED HOST memory and is streamed through the GPU: the page
    range is cut into chunks, chunk i+1 travels host->device on a copy stream into one of two staging
    buffers while the tensor-core kernel scores chunk i on the caller's stream, and every chunk's
    scores land directly in their columns of the (nq, np) result (the kernel takes a leading
    dimension).  For query batches large enough to be tensor-bound (about 900 queries at PCIe Gen5
    x16: 266 KB per page vs 8.4 MFLOP per pair) the copy hides behind the scoring, i.e. a corpus that
    does not fit -- or has not been loaded into -- HBM is scored at the resident rate.

    tokens_host: (np, L, 128) bf16, pinned, L % 16 == 0, every slot a valid token (filler slots hold a
    duplicate of one of the page's tokens, like PackedPages).  The reference keeps the corpus on the
    device (mainv2_iter_listwise.py:96); this class is the end-to-end "everything starts on the host"
    measurement of bench.py and the route for corpora beyond 180 GB per GPU.
    """

    def __init__(self, tokens_host: torch.Tensor, device, pages_per_chunk: int = 0, page_offset: int = 0):
        if tokens_host.is_cuda or not tokens_host.is_pinned():
            raise ValueError("HostCorpus needs a pinned host tensor")
        npg, L, d = tokens_host.shape
        if tokens_host.dtype != torch.bfloat16 or d != TC_DIM or L % CHUNK != 0 or not tokens_host.is_contiguous():
            raise ValueError("HostCorpus needs contiguous bf16 (np, L%16==0, 128)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EvdrError("HostCorpus scores on the GPU: this package has no CPU fallback")
        self.host = tokens_host
        self.n_pages, self.L = npg, L
        self.page_offset = int(page_offset)
        if pages_per_chunk <= 0:                       # ~2 GiB per staging buffer
            pages_per_chunk = max(1, (2 << 30) // (L * d * 2))
        self.ppc = min(int(pages_per_chunk), npg)
        self.ramp = 1.18
        self._stage = [torch.empty((self.ppc, L, d), dtype=torch.bfloat16, device=self.device) for _ in range(2)]
        cpp = L // CHUNK
        self._chunk_off = torch.arange(0, (self.ppc + 1) * cpp, cpp, dtype=torch.int32, device=self.device)
        self._copy_stream = torch.cuda.Stream(self.device)
        self._loaded = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]

    def score(self, Q: torch.Tensor, qmask: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """MaxSim of (nq, lq <= 128, 128) queries (device) against the host corpus -> fp32 (nq, np) on device."""
        Qs, ms, G = split_long_queries(Q, qmask)
        Qb, qm32 = pack_queries(Qs, ms)
        rows = Qb.shape[0]
        dev = self.device
        full = torch.empty((rows, self.n_pages), dtype=torch.float32, device=dev) if (out is None or G > 1) else out
        main = torch.cuda.current_stream(dev)
        cs = self._copy_stream
        # Chunk sizes ramp up from ppc/32 to ppc: only the FIRST copy is exposed, so it is kept small.  The
        # growth factor must stay below (link rate / scoring rate), else the copy of chunk i+1 outlasts the
        # scoring of chunk i all the way up the ramp (a 2x ramp exposed ~25 ms of a 0.6 s step): 1.18 covers
        # scoring down to ~47 GB/s of page bytes per second on a 55 GB/s link.
        bounds, c, step = [0], 0, float(max(1, self.ppc // 32))
        while c < self.n_pages:
            c = min(self.n_pages, c + int(step))
            bounds.append(c)
            step = min(float(self.ppc), step * self.ramp)
        # the staging buffers may still be read by a previous call on this stream
        cs.wait_stream(main)
        for i in range(len(bounds) - 1):
            b = i & 1
            c0, c1 = bounds[i], bounds[i + 1]
            with torch.cuda.stream(cs):
                if i >= 2:
                    cs.wait_event(self._consumed[b])       # the kernel of chunk i-2 is done with buffer b
                self._stage[b][: c1 - c0].copy_(self.host[c0:c1], non_blocking=True)
                self._loaded[b].record(cs)
            main.wait_event(self._loaded[b])
            n = c1 - c0
            pk = PackedPages(self._stage[b][:n].view(n * self.L, TC_DIM), self._chunk_off[: n + 1], n,
                             n * (self.L // CHUNK), self.page_offset + c0)
            pk.score_packed(Qb, qm32, out=full[:, c0:c1])
            self._consumed[b].record(main)
        res = group_sum(full, G)
        if out is not None and res is not out:
            out.copy_(res)
            return out
        return res

    def search(self, Q: torch.Tensor, qmask: torch.Tensor, k: int = 10):
        """Top-k pages per query over the host-resident corpus WITHOUT materialising the (nq, np) score matrix:
        every streamed chunk goes through the one-call search (scan with the top-k in its epilogue for up to 8
        queries), its k candidates per query are kept, and one merge over the chunks' candidates gives the result
        (scores (nq, k) fp32, ids (nq, k) int64 global page ids).  Same overlap of copies and scoring as `score`."""
        from .topk import topk_merge
        if Q.shape[1] > TC_LQ or k > TOPK_MAX:
            from .topk import topk
            return topk(self.score(Q, qmask), k, id_offset=self.page_offset)
        Qb, qm32 = pack_queries(Q, qmask)
        nq = Qb.shape[0]
        dev = self.device
        main = torch.cuda.current_stream(dev)
        cs = self._copy_stream
        bounds, c, step = [0], 0, float(max(1, self.ppc // 32))
        while c < self.n_pages:
            c = min(self.n_pages, c + int(step))
            bounds.append(c)
            step = min(float(self.ppc), step * self.ramp)
        n_chunks = len(bounds) - 1
        cand_s = torch.empty((nq, n_chunks, k), dtype=torch.float32, device=dev)
        cand_i = torch.empty((nq, n_chunks, k), dtype=torch.int64, device=dev)
        tmp_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        tmp_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        lib = _lib.load()
        h = _lib.handle(dev)
        cs.wait_stream(main)
        for i in range(n_chunks):
            b = i & 1
            c0, c1 = bounds[i], bounds[i + 1]
            with torch.cuda.stream(cs):
                if i >= 2:
                    cs.wait_event(self._consumed[b])
                self._stage[b][: c1 - c0].copy_(self.host[c0:c1], non_blocking=True)
                self._loaded[b].record(cs)
            main.wait_event(self._loaded[b])
            n = c1 - c0
            _lib.check(lib.evdr_sharded_search(h, Qb.data_ptr(), qm32.data_ptr(), nq, self._stage[b].data_ptr(),
                                               n * self.L, self._chunk_off.data_ptr(), n, n * (self.L // CHUNK), k,
                                               self.page_offset + c0, 0, tmp_s.data_ptr(), tmp_i.data_ptr(),
                                               _lib.stream_ptr(dev)), "evdr_sharded_search")
            cand_s[:, i].copy_(tmp_s)
            cand_i[:, i].copy_(tmp_i)
            self._consumed[b].record(main)
        return topk_merge(cand_s.view(nq, n_chunks * k), cand_i.view(nq, n_chunks * k), k)
