"""Top-k selection and trec_eval-style measures on device (include/evdr.h: evdr_topk,
evdr_topk_merge, evdr_retrieval_metrics).  Replaces the reference's dict build +
mteb/pytrec_eval tail (mainv2_iter_listwise.py:312-320, evaluator/retrieval.py:239-246).
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib
from ._lib import TOPK_MAX

METRIC_NAMES = ("ndcg", "map", "recall", "precision", "mrr")


def topk(scores: torch.Tensor, k: int, tie_rank: torch.Tensor | None = None, id_offset: int = 0):
    """Per row the k best columns by (score desc, tie_rank asc | column asc).

    Returns (top_scores (nq,k) fp32, top_ids (nq,k) int64); slots beyond the number of columns are
    (-inf, -1)."""
    _lib.require_cuda(scores, "scores")
    if scores.dtype != torch.float32 or scores.dim() != 2:
        raise ValueError("scores must be a 2-D float32 tensor")
    if scores.stride(1) != 1:
        scores = scores.contiguous()
    nq, npg = scores.shape
    if not (1 <= k <= TOPK_MAX):
        raise ValueError(f"k={k} outside [1,{TOPK_MAX}]")
    dev = scores.device
    if tie_rank is not None:
        tie_rank = tie_rank.to(device=dev, dtype=torch.int32).contiguous()
        if tie_rank.numel() != npg:
            raise ValueError("tie_rank must have one entry per column")
    top_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    top_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    lib = _lib.load()
    h = _lib.handle(dev)
    st = _lib.stream_ptr(dev)
    for q0 in range(0, nq, 65535):
        q1 = min(nq, q0 + 65535)
        _lib.check(lib.evdr_topk(h, scores[q0:q1].data_ptr(), scores.stride(0), q1 - q0, npg, k,
                                 _lib.ptr(tie_rank), id_offset, top_s[q0:q1].data_ptr(),
                                 top_i[q0:q1].data_ptr(), st), "evdr_topk")
    return top_s, top_i


def topk_merge(cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int):
    """Select the k best of explicit (score, id) candidates per row; ties by id ascending; id<0 skipped."""
    _lib.require_cuda(cand_scores, "cand_scores")
    cand_scores = cand_scores.to(torch.float32).contiguous()
    cand_ids = cand_ids.to(torch.int64).contiguous()
    nq, nc = cand_scores.shape
    dev = cand_scores.device
    top_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    top_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    lib = _lib.load()
    _lib.check(lib.evdr_topk_merge(_lib.handle(dev), cand_scores.data_ptr(), cand_ids.data_ptr(), nq, nc, k,
                                   top_s.data_ptr(), top_i.data_ptr(), _lib.stream_ptr(dev)),
               "evdr_topk_merge")
    return top_s, top_i


def retrieval_metrics(top_ids: torch.Tensor, qrel_off: torch.Tensor, qrel_ids: torch.Tensor,
                      qrel_rels: torch.Tensor, k_values: Sequence[int]) -> torch.Tensor:
    """Per-query measures from ranked ids and CSR qrels -> float64 (5, len(k_values), nq):
    ndcg, map, recall, precision, mrr (linear gain, log2 discount, as trec_eval)."""
    _lib.require_cuda(top_ids, "top_ids")
    dev = top_ids.device
    top_ids = top_ids.to(torch.int64).contiguous()
    nq, kmax = top_ids.shape
    qrel_off = qrel_off.to(device=dev, dtype=torch.int32).contiguous()
    qrel_ids = qrel_ids.to(device=dev, dtype=torch.int64).contiguous()
    qrel_rels = qrel_rels.to(device=dev, dtype=torch.int32).contiguous()
    if qrel_ids.numel() == 0:  # keep pointers valid
        qrel_ids = torch.zeros(1, dtype=torch.int64, device=dev)
        qrel_rels = torch.zeros(1, dtype=torch.int32, device=dev)
    kv = torch.tensor(list(k_values), dtype=torch.int32, device=dev)
    out = torch.empty((5, len(k_values), nq), dtype=torch.float64, device=dev)
    lib = _lib.load()
    _lib.check(lib.evdr_retrieval_metrics(_lib.handle(dev), top_ids.data_ptr(), kmax, nq, qrel_off.data_ptr(),
                                          qrel_ids.data_ptr(), qrel_rels.data_ptr(), kv.data_ptr(),
                                          len(k_values), out.data_ptr(), _lib.stream_ptr(dev)),
               "evdr_retrieval_metrics")
    return out


def topk_metrics(scores: torch.Tensor, kmax: int, qrel_off: torch.Tensor, qrel_ids: torch.Tensor,
                 qrel_rels: torch.Tensor, k_values: Sequence[int], tie_rank: torch.Tensor | None = None,
                 id_offset: int = 0):
    """Fused top-kmax + measures (evdr_topk_metrics): returns (top_scores, top_ids, metrics (5, nk, nq))."""
    _lib.require_cuda(scores, "scores")
    if scores.dtype != torch.float32 or scores.dim() != 2:
        raise ValueError("scores must be a 2-D float32 tensor")
    if scores.stride(1) != 1:
        scores = scores.contiguous()
    nq, npg = scores.shape
    if not (1 <= kmax <= TOPK_MAX):
        raise ValueError(f"kmax={kmax} outside [1,{TOPK_MAX}]")
    if nq > 65535:
        raise NotImplementedError("more than 65535 queries per call")
    dev = scores.device
    if tie_rank is not None:
        tie_rank = tie_rank.to(device=dev, dtype=torch.int32).contiguous()
    qrel_off = qrel_off.to(device=dev, dtype=torch.int32).contiguous()
    qrel_ids = qrel_ids.to(device=dev, dtype=torch.int64).contiguous()
    qrel_rels = qrel_rels.to(device=dev, dtype=torch.int32).contiguous()
    if qrel_ids.numel() == 0:
        qrel_ids = torch.zeros(1, dtype=torch.int64, device=dev)
        qrel_rels = torch.zeros(1, dtype=torch.int32, device=dev)
    kv = torch.tensor(list(k_values), dtype=torch.int32, device=dev)
    top_s = torch.empty((nq, kmax), dtype=torch.float32, device=dev)
    top_i = torch.empty((nq, kmax), dtype=torch.int64, device=dev)
    out = torch.empty((5, len(k_values), nq), dtype=torch.float64, device=dev)
    lib = _lib.load()
    _lib.check(lib.evdr_topk_metrics(_lib.handle(dev), scores.data_ptr(), scores.stride(0), nq, npg, kmax,
                                     _lib.ptr(tie_rank), id_offset, qrel_off.data_ptr(), qrel_ids.data_ptr(),
                                     qrel_rels.data_ptr(), kv.data_ptr(), len(k_values), top_s.data_ptr(),
                                     top_i.data_ptr(), out.data_ptr(), _lib.stream_ptr(dev)),
               "evdr_topk_metrics")
    return top_s, top_i, out
