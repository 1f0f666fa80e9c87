"""Host-side mirror of the reference's ``evaluator/retrieval.py`` for the MaxSim hot path.

Same names, argument meaning and error behaviour as the reference, so that its 21 main*.py
scripts run unchanged when this module is importable as ``evaluator.retrieval`` (see
``dropin/evaluator`` and INTEGRATION.md):

* ``score_multi_vector_masked(Q, P, qmask, pmask, chunk_p=128)``  <- evaluator/retrieval.py:166-213
* ``BaseVisualRetrieverProcessor.score_multi_vector / score_single_vector`` <- :78-150
* ``CustomRetrievalEvaluator.compute_mteb_metrics``               <- :220-255

All compute happens in libevdr.so (hand-written sm_100a CUDA) through its C ABI.  PyTorch only
owns the device memory and the stream.  Inputs on the CPU raise: there is no CPU fallback.
"""
from __future__ import annotations

import os
import threading
from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import _lib
from ._lib import TC_DIM, TC_LQ, TC_LQ_MAX
from .corpus import PackedPages, _float_operand, _mask_u8, pack_queries
from .topk import retrieval_metrics, topk, topk_metrics

__all__ = [
    "score_multi_vector_masked", "BaseVisualRetrieverProcessor", "CustomRetrievalEvaluator",
    "get_torch_device", "left_padding", "maxsim_exact", "TeacherScoreTable",
]

# rowmax scratch of the exact path is nq*lq*np floats; pages are processed in blocks that keep it
# below this many bytes (and below the 65535-page grid limit of one launch)
_EXACT_WS_BYTES = 1 << 30


def get_torch_device(device: str = "auto") -> str:
    """evaluator/retrieval.py:10-28 -- this package only ever answers with a CUDA device."""
    if device == "auto":
        if not torch.cuda.is_available():
            raise _lib.EvdrError("no CUDA device: this package has no CPU / MPS fallback")
        return "cuda:0"
    return device


def left_padding(sequences, batch_first=True, padding_value=0):
    """Zero LEFT padding of a list of (Li, D) tensors onto the GPU (evaluator/retrieval.py:30-45).
    Kept for API compatibility; the scorer below does not materialise padded blocks."""
    sequences = [s.unsqueeze(0) if s.ndim == 1 else s for s in sequences]
    max_len = max(s.size(0) for s in sequences)
    d = sequences[0].size(-1)
    out = torch.full((len(sequences), max_len, d), padding_value, device="cuda", dtype=sequences[0].dtype)
    for i, s in enumerate(sequences):
        out[i, max_len - s.size(0):] = s.to("cuda")
    return out if batch_first else out.transpose(0, 1)


# ---------------------------------------------------------------------------
# exact path (fp32 FMA, argmax, autograd)
# ---------------------------------------------------------------------------
def _exact_forward(Q, P, qm, pm, want_argmax: bool):
    """evdr_maxsim_fwd over page blocks.  Returns (out (nq,np) fp32, argmax int32 [np, nq*lq] | None)."""
    nq, lq, d = Q.shape
    npg, lp, _ = P.shape
    dev = Q.device
    lib = _lib.load()
    h = _lib.handle(dev)
    st = _lib.stream_ptr(dev)
    out = torch.empty((nq, npg), dtype=torch.float32, device=dev)
    R = nq * lq
    blk = max(1, min(65535, npg, _EXACT_WS_BYTES // (4 * R)))
    ws = torch.empty((blk, R), dtype=torch.float32, device=dev)
    am = torch.empty((npg, R), dtype=torch.int32, device=dev) if want_argmax else None
    qd, pd = _lib.dtype_code(Q), _lib.dtype_code(P)
    single = blk >= npg
    for c0 in range(0, npg, blk):
        c1 = min(npg, c0 + blk)
        o = out if single else torch.empty((nq, c1 - c0), dtype=torch.float32, device=dev)
        _lib.check(lib.evdr_maxsim_fwd(h, Q.data_ptr(), qd, P[c0:c1].data_ptr(), pd, qm.data_ptr(),
                                       pm[c0:c1].data_ptr(), nq, c1 - c0, lq, lp, d, ws.data_ptr(),
                                       None if am is None else am[c0:c1].data_ptr(), o.data_ptr(), st),
                   "evdr_maxsim_fwd")
        if not single:
            out[:, c0:c1] = o
    return out, am


class _MaxSimExactFn(torch.autograd.Function):
    """A1 + A5: exact forward with argmax, deterministic closed-form backward."""

    @staticmethod
    def forward(ctx, Q, P, qm, pm):
        out, am = _exact_forward(Q, P, qm, pm, want_argmax=True)
        ctx.save_for_backward(Q, P, qm, pm, am)
        return out

    @staticmethod
    def backward(ctx, g):
        return _backward_through_argmax(ctx, g)


def _backward_through_argmax(ctx, g):
    Q, P, qm, pm, am = ctx.saved_tensors
    need_q, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    nq, lq, d = Q.shape
    npg, lp, _ = P.shape
    dev = Q.device
    g = g.to(torch.float32).contiguous()
    dP = torch.empty((npg, lp, d), dtype=torch.float32, device=dev) if need_p else None
    dQ = torch.empty((nq, lq, d), dtype=torch.float32, device=dev) if need_q else None
    if need_p or need_q:
        lib = _lib.load()
        _lib.check(lib.evdr_maxsim_bwd(_lib.handle(dev), g.data_ptr(), Q.data_ptr(), _lib.dtype_code(Q),
                                       P.data_ptr(), _lib.dtype_code(P), qm.data_ptr(), pm.data_ptr(),
                                       am.data_ptr(), nq, npg, lq, lp, d, _lib.ptr(dP), _lib.ptr(dQ),
                                       _lib.stream_ptr(dev)), "evdr_maxsim_bwd")
    return (None if dQ is None else dQ.to(Q.dtype), None if dP is None else dP.to(P.dtype), None, None)


class _MaxSimRescoreFn(torch.autograd.Function):
    """Gradient calls on ColPali shapes: ONE tcgen05 pass on bf16x3-split operands gives the scores (to
    ~5e-6 relative of the fp32 values) and the winning page token of every (query token, page); the
    backward scatters through those indices.  EVDR_TRAIN_RESCORE=1 recomputes the winners' dot products
    exactly in fp32 on the original operands instead (evdr_maxsim_rescore), as does the composed path for
    problems beyond one call."""

    @staticmethod
    def forward(ctx, Q, P, qm, pm):
        nq, lq, d = Q.shape
        npg, lp, _ = P.shape
        dev = Q.device
        lib = _lib.load()
        h = _lib.handle(dev)
        st = _lib.stream_ptr(dev)
        R = nq * lq
        if 4 * R * npg <= _EXACT_WS_BYTES and npg <= 65535:
            # one library call (evdr_maxsim_train_fwd): both packs, the search, the rescoring and the sum
            am = torch.empty((npg, R), dtype=torch.int32, device=dev)
            out = torch.empty((nq, npg), dtype=torch.float32, device=dev)
            rescore = 1 if os.environ.get("EVDR_TRAIN_RESCORE", "0") == "1" else 0
            _lib.check(lib.evdr_maxsim_train_fwd(h, Q.data_ptr(), _lib.dtype_code(Q), P.data_ptr(), _lib.dtype_code(P),
                                                 qm.data_ptr(), pm.data_ptr(), nq, npg, lq, lp, d, rescore,
                                                 am.data_ptr(), out.data_ptr(), st), "evdr_maxsim_train_fwd")
            ctx.save_for_backward(Q, P, qm, pm, am)
            return out
        # larger problems: the same steps composed here, the rescoring in page blocks
        # bf16x3 split operands: the search runs at ~fp32 accuracy, so the winners are the reference's.
        # Queries of 33..128 tokens arrive zero-padded to a multiple of 32 (score_multi_vector_masked) and
        # are searched as nq*G rows of 32 tokens: row-major (nq, G*32) IS (nq*G, 32), so the argmax table
        # [np, nq*lq] needs no re-indexing
        G = 1
        if lq > TC_LQ:
            assert lq % TC_LQ == 0, "long queries must be padded to a multiple of 32 tokens"
            G = lq // TC_LQ
        Qb, qm32 = pack_queries(Q.view(nq * G, lq // G, d), qm.view(nq * G, lq // G), 3)
        pk = PackedPages.from_padded(P, pm, normalize=False, varlen=False, k_blocks=3)
        R = nq * lq
        am = torch.empty((npg, R), dtype=torch.int32, device=dev)
        _lib.check(lib.evdr_maxsim_packed_argmax(h, Qb.data_ptr(), qm32.data_ptr(), nq * G, lq // G, pk.tokens.data_ptr(),
                                                 pk.tokens.shape[0], pk.chunk_off.data_ptr(), npg,
                                                 pk.total_chunks, 3, am.data_ptr(), st), "evdr_maxsim_packed_argmax")
        out = torch.empty((nq, npg), dtype=torch.float32, device=dev)
        blk = max(1, min(65535, npg, _EXACT_WS_BYTES // (4 * R)))
        ws = torch.empty((blk, R), dtype=torch.float32, device=dev)
        single = blk >= npg
        qd, pd = _lib.dtype_code(Q), _lib.dtype_code(P)
        for c0 in range(0, npg, blk):
            c1 = min(npg, c0 + blk)
            o = out if single else torch.empty((nq, c1 - c0), dtype=torch.float32, device=dev)
            _lib.check(lib.evdr_maxsim_rescore(h, Q.data_ptr(), qd, P[c0:c1].data_ptr(), pd, qm.data_ptr(),
                                               pm[c0:c1].data_ptr(), am[c0:c1].data_ptr(), nq, c1 - c0, lq, lp,
                                               d, ws.data_ptr(), o.data_ptr(), st), "evdr_maxsim_rescore")
            if not single:
                out[:, c0:c1] = o
        ctx.save_for_backward(Q, P, qm, pm, am)
        return out

    @staticmethod
    def backward(ctx, g):
        return _backward_through_argmax(ctx, g)


def maxsim_exact(Q, P, qmask, pmask) -> torch.Tensor:
    """fp32-exact MaxSim (any lq, lp, d % 4 == 0); differentiable w.r.t. P and Q."""
    _lib.require_cuda(Q, "Q")
    _lib.require_cuda(P, "P")
    Qc, Pc = _float_operand(Q), _float_operand(P)
    qm, pm = _mask_u8(qmask), _mask_u8(pmask)
    if Qc.shape[-1] % 4 != 0:
        raise NotImplementedError("embedding dim must be a multiple of 4")
    if torch.is_grad_enabled() and (Qc.requires_grad or Pc.requires_grad):
        return _MaxSimExactFn.apply(Qc, Pc, qm, pm)
    return _exact_forward(Qc.detach(), Pc.detach(), qm, pm, want_argmax=False)[0]


# ---------------------------------------------------------------------------
# packed-corpus cache for frozen page tensors (the teacher is scored every step with the same P,
# mainv2_iter_listwise.py:283-284); keyed on storage + in-place version counters
# ---------------------------------------------------------------------------
_PACK_CACHE: "OrderedDict[tuple, PackedPages]" = OrderedDict()
_PACK_CACHE_SLOTS = int(os.environ.get("EVDR_PACK_CACHE", "2"))
_PACK_CACHE_LOCK = threading.Lock()


def default_precision(P: torch.Tensor) -> str:
    """Arithmetic of calls that do not name a precision.  The reference up-casts to fp32 and runs an fp32
    einsum (evaluator/retrieval.py:176-177,195).  A bf16 page tensor is scored as plain bf16 x bf16 -> fp32
    (its values ARE bf16, only the query is rounded); an fp32 page tensor (the teacher / student tensors of
    the training scripts) is scored with the bf16x3 split, i.e. the fp32 dot products to ~2^-18 relative, so
    that teacher and student scores of one step -- and evaluation rankings -- carry the same, near-fp32
    accuracy.  EVDR_PRECISION=bf16|bf16x3|fp32 overrides."""
    env = os.environ.get("EVDR_PRECISION")
    if env:
        return env
    return "bf16" if P.dtype == torch.bfloat16 else "bf16x3"


def _packed_for(P: torch.Tensor, pmask: torch.Tensor, k_blocks: int = 1) -> PackedPages:
    """Packed stream of a page tensor, cached for frozen tensors.  The key is storage + in-place version
    counters, so writes through torch ops invalidate an entry; writes that bypass the counters (``.data``,
    raw pointers) are not seen -- set EVDR_PACK_CACHE=0 for such callers.  Inference-mode tensors carry no
    version counter and are never cached."""
    cacheable = _PACK_CACHE_SLOTS > 0 and not (P.is_inference() or pmask.is_inference())
    key = None
    if cacheable:
        key = (P.data_ptr(), P._version, tuple(P.shape), P.dtype, pmask.data_ptr(), pmask._version,
               tuple(pmask.shape), pmask.dtype, P.device.index, k_blocks)
        with _PACK_CACHE_LOCK:
            hit = _PACK_CACHE.get(key)
            if hit is not None:
                _PACK_CACHE.move_to_end(key)
                return hit
    if (k_blocks == 1 and P.dtype == torch.bfloat16 and P.is_contiguous() and P.shape[1] % _lib.CHUNK == 0
            and bool(pmask.all().item())):          # one host sync per cache MISS (frozen corpora: once)
        pk = PackedPages.from_uniform_tokens(P)        # already in stream layout: used in place
    else:
        pk = PackedPages.from_padded(P, pmask, normalize=False, varlen=False, k_blocks=k_blocks)
    if cacheable:
        # hold references so that the key's storage cannot be recycled while cached
        pk._key_refs = (P, pmask)
        with _PACK_CACHE_LOCK:
            _PACK_CACHE[key] = pk
            while len(_PACK_CACHE) > _PACK_CACHE_SLOTS:
                _PACK_CACHE.popitem(last=False)
    return pk


def clear_pack_cache() -> None:
    _PACK_CACHE.clear()


class _ScoreMatrix(torch.Tensor):
    """The (Q, P) fp32 score tensor of a no-grad call: an ordinary CUDA tensor for every torch op (ops on
    it return plain tensors, like nn.Parameter), except that SCALAR indexing ``scores[qi, di]`` is served
    from a pinned host mirror that is filled with one device->host copy on first use.

    Why: every evaluation of the reference builds its results dict with ``float(scores[qi, di].item())``
    for all Q x P pairs (mainv2_iter_listwise.py:312-318) -- on a GPU tensor that is Q*P device
    synchronisations (250 000 for DocVQA, seconds per evaluation).  With the mirror the
    same unmodified loop does one copy and Q*P host reads.  The mirror is keyed on the tensor's in-place
    version counter, so writes through any torch op invalidate it."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    @staticmethod
    def wrap(t: torch.Tensor) -> "torch.Tensor":
        if os.environ.get("EVDR_HOST_MIRROR", "1") == "0":
            return t
        out = torch.Tensor._make_subclass(_ScoreMatrix, t)
        out._evdr_mirror = None
        return out

    def _host(self):
        m = getattr(self, "_evdr_mirror", None)
        if m is None or m[0] != self._version:
            plain = torch.Tensor._make_subclass(torch.Tensor, self)
            host = torch.empty(plain.shape, dtype=plain.dtype, pin_memory=True)
            host.copy_(plain)                       # synchronous for pinned destinations: one sync in total
            m = (self._version, host)
            self._evdr_mirror = m
        return m[1]

    def __getitem__(self, idx):
        if (type(idx) is tuple and len(idx) == self.dim() and all(type(i) is int for i in idx)):
            return self._host()[idx]
        return torch.Tensor._make_subclass(torch.Tensor, self)[idx]


def score_multi_vector_masked(
    Q: torch.Tensor,        # (Q, Lq, D)
    P: torch.Tensor,        # (P, Lp, D)
    qmask: torch.Tensor,    # (Q, Lq) any dtype, non-zero = valid
    pmask: torch.Tensor,    # (P, Lp)
    chunk_p: int = 128,
    *,
    precision: Optional[str] = None,
) -> torch.Tensor:
    """Masked MaxSim, drop-in for evaluator/retrieval.py:166-213.

        out[q,c] = has[c] * sum_n qmask[q,n] * max_m ( pmask[c,m] ? <Q[q,n],P[c,m]> : -1e4 )

    Returns a new fp32 (Q, P) tensor on the inputs' device.  ``chunk_p`` is accepted and ignored
    (the reference's outputs do not depend on it and nothing is materialised here).  Results that carry
    no gradient are ``_ScoreMatrix`` tensors: plain CUDA tensors whose scalar reads ``scores[q, c].item()``
    (the reference's evaluation loop, mainv2_iter_listwise.py:312-318) come from a host mirror filled by
    ONE device->host copy instead of Q*P synchronisations (EVDR_HOST_MIRROR=0 disables it).

    precision=None: ``default_precision(P)`` -- plain bf16 operands for a bf16 page tensor, the bf16x3 split
    (near-fp32 similarities) for an fp32 one; both run on the tcgen05 kernel whenever D == 128 and
    Lq <= 128.  Calls that need a gradient run ONE bf16x3 pass that yields the scores and the winning page
    token of every (query token, page); the backward is the deterministic closed form of the reference's
    autograd through those indices (EVDR_TRAIN_RESCORE=1 additionally rescores the winners in fp32 on the
    original operands).  precision="fp32" (or env EVDR_PRECISION=fp32) and other shapes use the exact fp32
    kernels throughout.
    """
    if Q.dim() != 3 or P.dim() != 3 or Q.shape[-1] != P.shape[-1]:
        raise ValueError(f"expected Q (Q,Lq,D) and P (P,Lp,D), got {tuple(Q.shape)} and {tuple(P.shape)}")
    if qmask.shape != Q.shape[:2] or pmask.shape != P.shape[:2]:
        raise ValueError("mask shapes must match the leading two dims of Q / P")
    _lib.require_cuda(Q, "Q")
    _lib.require_cuda(P, "P")
    if Q.device != P.device:
        raise ValueError("Q and P must be on the same device")
    nq, lq, d = Q.shape
    npg = P.shape[0]
    if nq == 0 or npg == 0:
        return torch.zeros((nq, npg), dtype=torch.float32, device=Q.device)
    precision = precision or default_precision(P)
    needs_grad = torch.is_grad_enabled() and (Q.requires_grad or P.requires_grad)
    if precision == "fp32" or d != TC_DIM or lq > TC_LQ_MAX:
        out = maxsim_exact(Q, P, qmask, pmask)
        return out if out.requires_grad else _ScoreMatrix.wrap(out)
    if precision == "bf16x3" and not needs_grad:
        # high-fidelity tensor-core mode for operands that are not bf16-representable
        with torch.no_grad():
            return _ScoreMatrix.wrap(_packed_for(P.detach(), pmask, 3).score(Q.detach(), qmask))
    if needs_grad:
        # training: tensor-core search + fp32 rescoring of the winners, backward through the same indices
        Qf, qm8 = _float_operand(Q), _mask_u8(qmask)
        if lq > TC_LQ and lq % TC_LQ:
            # 33..128-token queries: masked zero tokens up to a multiple of 32 (they add 0 to every score
            # and receive no gradient; autograd slices dQ back)
            pad = TC_LQ - lq % TC_LQ
            Qf = torch.nn.functional.pad(Qf, (0, 0, 0, pad))
            qm8 = torch.nn.functional.pad(qm8, (0, pad))
        return _MaxSimRescoreFn.apply(Qf, _float_operand(P), qm8, _mask_u8(pmask))
    with torch.no_grad():
        return _ScoreMatrix.wrap(_packed_for(P.detach(), pmask).score(Q.detach(), qmask))   # lq > 32: 32-token groups


class TeacherScoreTable:
    """One-shot teacher scores.  The teacher pages are frozen and the pseudo-query
    set is fixed, yet every step of the reference recomputes ``score_multi_vector_masked(Qb,
    P_teacher_norm, qmb, pmask_teacher)`` under no_grad (mainv2_iter_listwise.py:283-284) -- each row
    30 times over 30 epochs, always on the expensive 1030-token pages.  Scoring ALL queries against
    the teacher once is a single tensor-core-bound call (25 000 queries x 1000 pages: ~0.15 s) whose
    (Nq, Np) fp32 table is 100 MB; a training step then gathers its rows by dataset index.

        table = TeacherScoreTable(Q_all, qmask_all, P_teacher_norm, pmask_teacher)
        sc_t = table[batch_indices]          # (B, Np) fp32, == score_multi_vector_masked(Q_all[idx], ...)
    """

    def __init__(self, Q_all: torch.Tensor, qmask_all: torch.Tensor, P_teacher: torch.Tensor,
                 pmask_teacher: torch.Tensor, query_block: int = 8192):
        with torch.no_grad():
            rows = [score_multi_vector_masked(Q_all[i:i + query_block], P_teacher, qmask_all[i:i + query_block],
                                              pmask_teacher) for i in range(0, Q_all.shape[0], query_block)]
        self.scores = rows[0] if len(rows) == 1 else torch.cat(rows, dim=0)

    def __getitem__(self, idx) -> torch.Tensor:
        idx = torch.as_tensor(idx, device=self.scores.device)
        return self.scores.index_select(0, idx.reshape(-1).to(torch.int64))

    def __len__(self) -> int:
        return self.scores.shape[0]


# ---------------------------------------------------------------------------
# list-input scorer (A2)
# ---------------------------------------------------------------------------
class BaseVisualRetrieverProcessor(ABC):
    """Mirror of evaluator/retrieval.py:47-164 (scoring statics only carry compute)."""

    @abstractmethod
    def process_images(self, images):
        pass

    @abstractmethod
    def process_queries(self, queries, max_length: int = 50, suffix: Optional[str] = None):
        pass

    @abstractmethod
    def score(self, qs, ps, device=None, **kwargs) -> torch.Tensor:
        pass

    @staticmethod
    def score_single_vector(qs: List[torch.Tensor], ps: List[torch.Tensor],
                            device: Optional[Union[str, torch.device]] = None) -> torch.Tensor:
        """Dense dot scores (evaluator/retrieval.py:78-99): a plain library GEMM, not MaxSim."""
        device = device or get_torch_device("auto")
        if len(qs) == 0:
            raise ValueError("No queries provided")
        if len(ps) == 0:
            raise ValueError("No passages provided")
        q = torch.stack(list(qs)).to(device)
        p = torch.stack(list(ps)).to(device)
        return (q.to(torch.float32) @ p.to(torch.float32).T).to(torch.float32)

    @staticmethod
    def score_multi_vector(
        qs: Union[torch.Tensor, List[torch.Tensor]],
        ps: Union[torch.Tensor, List[torch.Tensor]],
        batch_size: int = 128,
        device: Optional[Union[str, torch.device]] = None,
    ) -> torch.Tensor:
        """List-input MaxSim without masks (evaluator/retrieval.py:101-150).

        The reference zero-LEFT-pads every block of ``batch_size`` pages to the block's longest page
        and lets the padded zero tokens take part in the max, so a page shorter than its block's
        longest gets max(true max, 0) per query token.  Here such a page simply receives one extra
        all-zero token; nothing is padded and the result is identical.  Returns fp32 on the CPU like
        the reference (:141,149).
        """
        device = torch.device(device or get_torch_device("auto"))
        if len(qs) == 0:
            raise ValueError("No queries provided")
        if len(ps) == 0:
            raise ValueError("No passages provided")
        qs = [q.unsqueeze(0) if q.ndim == 1 else q for q in qs]
        ps = [p.unsqueeze(0) if p.ndim == 1 else p for p in ps]
        d = qs[0].shape[-1]
        qlens = [q.shape[0] for q in qs]
        lq = max(qlens)
        # queries: one pad_sequence + one copy (zero RIGHT padding with a mask == the reference's zero left
        # padding: padded query tokens add 0 either way)
        if len({q.device for q in qs}) > 1:            # mixed devices: the reference moves each one (:38-43)
            qs = [q.to(device) for q in qs]
        Q = torch.nn.utils.rnn.pad_sequence([q.to(torch.float32) for q in qs], batch_first=True).to(device)
        qm = (torch.arange(lq)[None, :] < torch.tensor(qlens)[:, None]).to(device)
        lens = [p.shape[0] for p in ps]
        padded = []
        for j0 in range(0, len(ps), batch_size):
            blk = lens[j0:j0 + batch_size]
            mx = max(blk)
            padded.extend(l < mx for l in blk)
        with torch.no_grad():
            prec = os.environ.get("EVDR_PRECISION") or (
                "bf16" if all(p.dtype == torch.bfloat16 for p in ps) else "bf16x3")     # see default_precision
            if d == TC_DIM and lq <= TC_LQ_MAX and prec != "fp32":
                # pages: one concatenation, one copy, one ragged pack; a page shorter than its block's longest
                # gets one extra all-zero token (the reference's padding zeros take part in the max)
                pk = PackedPages.from_tensor_list(ps, device=device, extra_zero_token=padded,
                                                  k_blocks=3 if prec == "bf16x3" else 1)
                scores = pk.score(Q, qm)
            else:
                lp = max(l + (1 if pad else 0) for l, pad in zip(lens, padded))
                if len({p.device for p in ps}) > 1:
                    ps = [p.to(device) for p in ps]
                P = torch.nn.utils.rnn.pad_sequence([p.to(torch.float32) for p in ps], batch_first=True)
                if P.shape[1] < lp:
                    P = torch.nn.functional.pad(P, (0, 0, 0, lp - P.shape[1]))
                P = P.to(device)
                plen = torch.tensor([l + (1 if pad else 0) for l, pad in zip(lens, padded)])
                pm = (torch.arange(lp)[None, :] < plen[:, None]).to(device)   # the extra valid slot is a zero token
                scores = maxsim_exact(Q, P, qm, pm)
        scores = scores.cpu()
        assert scores.shape[0] == len(qs), f"Expected {len(qs)} scores, got {scores.shape[0]}"
        return scores.to(torch.float32)

    @abstractmethod
    def get_n_patches(self, image_size: Tuple[int, int], patch_size: int = 14, *args, **kwargs):
        pass


# ---------------------------------------------------------------------------
# metrics (A7)
# ---------------------------------------------------------------------------
class CustomRetrievalEvaluator:
    """Mirror of evaluator/retrieval.py:220-255.  The reference hands the dicts to mteb's
    RetrievalEvaluator (pytrec_eval); here the ranking (top-k with trec_eval's tie rule) and the
    measures run on the GPU, the host only converts dicts <-> arrays and takes mean + round(5).
    """

    def __init__(self, k_values: List[int] = [1, 3, 5, 10, 50, 70, 100], score_function: str = "cos_sim"):
        self.k_values = list(k_values)
        self.score_function = score_function

    # fast path: scores already on the device -------------------------------------------------
    def evaluate_scores(self, scores: torch.Tensor, qrel_off, qrel_ids, qrel_rels,
                        tie_rank: Optional[torch.Tensor] = None,
                        n_judged: Optional[int] = None) -> Dict[str, Dict[str, float]]:
        """scores (nq, np) fp32 on device; qrels as CSR over column indices.  ``n_judged`` is the MRR
        denominator -- ``len(relevant_docs)`` of the dict signature (mteb's mrr divides by the number of
        JUDGED queries, scored or not); it defaults to the number of scored rows."""
        kmax = min(max(self.k_values), _lib.TOPK_MAX)
        if max(self.k_values) > _lib.TOPK_MAX:
            raise NotImplementedError(f"k_values above {_lib.TOPK_MAX} are not supported")
        _, ids, m = topk_metrics(scores, kmax, qrel_off, qrel_ids, qrel_rels, self.k_values, tie_rank=tie_rank)
        if tie_rank is not None:  # mteb's MRR sorts stably: ties in column order
            _, ids2 = topk(scores, kmax, tie_rank=None)
            m[4] = retrieval_metrics(ids2, qrel_off, qrel_ids, qrel_rels, self.k_values)[4]
        return self._format(m.cpu(), mrr_den=n_judged)

    def _format(self, m: torch.Tensor, mrr_den: Optional[int] = None) -> Dict[str, Dict[str, float]]:
        def mean(row, den=None):  # Python float sum like mteb (sum(list)/len)
            vals = row.tolist()
            den = len(vals) if den is None else den
            return round(sum(vals) / den, 5) if vals and den else 0.0

        ks = self.k_values
        return {
            "NDCG": {f"NDCG@{k}": mean(m[0, j]) for j, k in enumerate(ks)},
            "mAP": {f"MAP@{k}": mean(m[1, j]) for j, k in enumerate(ks)},
            "Recall": {f"Recall@{k}": mean(m[2, j]) for j, k in enumerate(ks)},
            "Precision": {f"P@{k}": mean(m[3, j]) for j, k in enumerate(ks)},
            # mteb's mrr divides by len(qrels): judged queries without results count as 0
            "mRR": {f"MRR@{k}": mean(m[4, j], mrr_den) for j, k in enumerate(ks)},
        }

    # reference signature ------------------------------------------------------------------------
    def compute_mteb_metrics(self, relevant_docs: Dict[str, Dict[str, int]],
                             results: Dict[str, Dict[str, float]], **kwargs) -> Dict[str, Any]:
        """Same inputs/outputs as evaluator/retrieval.py:230-255."""
        if kwargs.get("ignore_identical_ids", False):
            results = {q: {d: s for d, s in ds.items() if d != q} for q, ds in results.items()}
        qids = [q for q in results if q in relevant_docs]
        if not qids:
            empty = torch.zeros((5, len(self.k_values), 0), dtype=torch.float64)
            return self._format(empty)
        # document universe in first-seen order (mteb's MRR is stable in dict order)
        doc_index: Dict[str, int] = {}
        for q in qids:
            for dname in results[q]:
                if dname not in doc_index:
                    doc_index[dname] = len(doc_index)
        nd = len(doc_index)
        dev = torch.device("cuda", torch.cuda.current_device())
        dense = torch.full((len(qids), nd), float("-inf"), dtype=torch.float32)
        first = list(results[qids[0]].keys())
        for i, q in enumerate(qids):
            ds = results[q]
            if len(ds) == nd and list(ds.keys()) == first:
                dense[i] = torch.tensor(list(ds.values()), dtype=torch.float32)
            else:
                idx = torch.tensor([doc_index[dname] for dname in ds], dtype=torch.long)
                dense[i, idx] = torch.tensor(list(ds.values()), dtype=torch.float32)
        # trec_eval breaks score ties by docid string, descending
        order = sorted(doc_index, reverse=True)
        tie = torch.empty(nd, dtype=torch.int32)
        tie[torch.tensor([doc_index[dname] for dname in order], dtype=torch.long)] = torch.arange(nd, dtype=torch.int32)
        off, ids, rels = [0], [], []
        extra = nd
        for q in qids:
            for dname, r in relevant_docs[q].items():
                j = doc_index.get(dname)
                if j is None:            # judged but never retrieved: still counts in recall / ideal DCG
                    j = extra
                    extra += 1
                ids.append(j)
                rels.append(int(r))
            off.append(len(ids))
        scores = dense.to(dev)
        # pages absent from a query's result list must never be ranked
        kmax = min(max(self.k_values), _lib.TOPK_MAX)
        if max(self.k_values) > _lib.TOPK_MAX:
            raise NotImplementedError(f"k_values above {_lib.TOPK_MAX} are not supported")
        tie_d = tie.to(dev)
        s1, i1 = topk(scores, kmax, tie_rank=tie_d)
        i1 = torch.where(torch.isinf(s1) & (s1 < 0), torch.full_like(i1, -1), i1)
        s2, i2 = topk(scores, kmax, tie_rank=None)
        i2 = torch.where(torch.isinf(s2) & (s2 < 0), torch.full_like(i2, -1), i2)
        off_t = torch.tensor(off, dtype=torch.int32)
        ids_t = torch.tensor(ids, dtype=torch.int64)
        rels_t = torch.tensor(rels, dtype=torch.int32)
        m = retrieval_metrics(i1, off_t, ids_t, rels_t, self.k_values)
        m[4] = retrieval_metrics(i2, off_t, ids_t, rels_t, self.k_values)[4]
        return self._format(m.cpu(), mrr_den=len(relevant_docs))
