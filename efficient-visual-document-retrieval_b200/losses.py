"""Score-matrix consumers behind the scorer, mirroring the names of the reference's ``criterion.py``
(SURVEY.md 8a A6 / 8f row 4).  Each call is ONE launch in libevdr.so that produces the loss value(s)
and the gradient w.r.t. the student scores -- the ``g`` that evdr_maxsim_bwd consumes.

* ``listwise_distillation_loss(score_s, score_t, k=10, temperature=1.0)``   <- criterion.py:113-142
* ``score_preserving_loss(score_s, score_t)``                               <- criterion.py:73-83
* ``liscore_losses(score_s, score_t, k, temperature)`` -> both of the above from one launch (the
  pair every ``*_liscore`` recipe adds up, mainv2_iter_liscore.py:301-303)
* ``pairwise_distillation_loss(score_s, score_t)``                          <- criterion.py:89-110
  without the three (B, N, N) intermediates
* ``rank_gap(sc_t, sc_s, k, aux_docs)``   <- mainv3_iter_liscore_QA_hardtoken.py:368-385

* ``lambda_loss(score_s, score_t, alpha, eps)``                             <- criterion.py:148-189
* ``ranknce_loss(score_s, score_t, temperature, lambda_weight)``            <- criterion.py:192-226
* ``infonce_supervised_loss`` / ``infonce_distillation_loss``               <- criterion.py:43-68 (one
  library cross-entropy call each; listed so that ``dropin/criterion.py`` covers the whole module)

The teacher scores are treated as constants exactly like the reference's ``score_t.detach()``.
"""
from __future__ import annotations

from typing import NamedTuple

import torch

from . import _lib


def _rows(x: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(x, name)
    if x.dim() != 2:
        raise ValueError(f"{name} must be (B, N), got {tuple(x.shape)}")
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    if x.stride(1) != 1:
        x = x.contiguous()
    return x


class _LiscoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score_s, score_t, k, temperature):
        s, t = _rows(score_s, "score_s"), _rows(score_t, "score_t")
        if s.shape != t.shape:
            raise ValueError(f"score_s {tuple(s.shape)} and score_t {tuple(t.shape)} differ")
        B, n = s.shape
        dev = s.device
        g_list = torch.empty((B, n), dtype=torch.float32, device=dev)
        g_mse = torch.empty((B, n), dtype=torch.float32, device=dev)
        ws = torch.empty(2 * B, dtype=torch.float32, device=dev)
        loss = torch.empty(2, dtype=torch.float32, device=dev)
        lib = _lib.load()
        _lib.check(lib.evdr_liscore_loss(_lib.handle(dev), s.data_ptr(), s.stride(0), t.data_ptr(), t.stride(0), B, n,
                                         int(k), float(temperature), g_list.data_ptr(), g_mse.data_ptr(),
                                         ws.data_ptr(), loss.data_ptr(), _lib.stream_ptr(dev)), "evdr_liscore_loss")
        ctx.save_for_backward(g_list, g_mse)
        ctx.in_dtype = score_s.dtype
        return loss[0], loss[1]

    @staticmethod
    def backward(ctx, gl, gm):
        g_list, g_mse = ctx.saved_tensors
        g = gl * g_list + gm * g_mse
        return g.to(ctx.in_dtype), None, None, None


def liscore_losses(score_s: torch.Tensor, score_t: torch.Tensor, k: int = 10, temperature: float = 1.0):
    """(listwise_distillation_loss, score_preserving_loss) of criterion.py from one launch."""
    return _LiscoreFn.apply(score_s, score_t, k, temperature)


def listwise_distillation_loss(score_s: torch.Tensor, score_t: torch.Tensor, k: int = 10,
                               temperature: float = 1.0) -> torch.Tensor:
    """criterion.py:113-142: -sum_{j in teacher top-k} softmax(t/T)_j * log_softmax(s/T)_j, mean over
    queries, times T^2.  Raises ValueError when k > N (torch.topk raises there)."""
    return _LiscoreFn.apply(score_s, score_t, k, temperature)[0]


def score_preserving_loss(score_s: torch.Tensor, score_t: torch.Tensor) -> torch.Tensor:
    """criterion.py:73-83: F.mse_loss(score_s, score_t.detach())."""
    return _LiscoreFn.apply(score_s, score_t, 1, 1.0)[1]


class _PairwiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score_s, score_t):
        s, t = _rows(score_s, "score_s"), _rows(score_t, "score_t")
        if s.shape != t.shape:
            raise ValueError(f"score_s {tuple(s.shape)} and score_t {tuple(t.shape)} differ")
        B, n = s.shape
        dev = s.device
        grad = torch.empty((B, n), dtype=torch.float32, device=dev)
        ws = torch.empty(B * ((n + 255) // 256), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        lib = _lib.load()
        _lib.check(lib.evdr_pairwise_loss(_lib.handle(dev), s.data_ptr(), s.stride(0), t.data_ptr(), t.stride(0), B, n,
                                          grad.data_ptr(), ws.data_ptr(), loss.data_ptr(), _lib.stream_ptr(dev)),
                   "evdr_pairwise_loss")
        ctx.save_for_backward(grad)
        ctx.in_dtype = score_s.dtype
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (g * grad).to(ctx.in_dtype), None


def pairwise_distillation_loss(score_s: torch.Tensor, score_t: torch.Tensor) -> torch.Tensor:
    """criterion.py:89-110: mean over (b, i, j) of BCE-with-logits(s_i - s_j, sigmoid(t_i - t_j))."""
    return _PairwiseFn.apply(score_s, score_t)


class _GradLossFn(torch.autograd.Function):
    """Shared shape of the single-output losses: the kernel returns the loss and d(loss)/d(score_s)."""

    @staticmethod
    def forward(ctx, score_s, score_t, which, p0, p1):
        s, t = _rows(score_s, "score_s"), _rows(score_t, "score_t")
        if s.shape != t.shape:
            raise ValueError(f"score_s {tuple(s.shape)} and score_t {tuple(t.shape)} differ")
        B, n = s.shape
        dev = s.device
        grad = torch.empty((B, n), dtype=torch.float32, device=dev)
        lib = _lib.load()
        h, st = _lib.handle(dev), _lib.stream_ptr(dev)
        if which == "lambda":
            ws = torch.empty(B * ((n + 255) // 256), dtype=torch.float32, device=dev)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            _lib.check(lib.evdr_lambda_loss(h, s.data_ptr(), s.stride(0), t.data_ptr(), t.stride(0), B, n, float(p0),
                                            float(p1), grad.data_ptr(), ws.data_ptr(), loss.data_ptr(), st),
                       "evdr_lambda_loss")
        else:
            ws = torch.empty(2 * B, dtype=torch.float32, device=dev)
            loss = torch.empty(3, dtype=torch.float32, device=dev)
            _lib.check(lib.evdr_ranknce_loss(h, s.data_ptr(), s.stride(0), t.data_ptr(), t.stride(0), B, n, float(p0),
                                             float(p1), grad.data_ptr(), ws.data_ptr(), loss.data_ptr(), st),
                       "evdr_ranknce_loss")
        ctx.save_for_backward(grad)
        ctx.in_dtype = score_s.dtype
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (g * grad).to(ctx.in_dtype), None, None, None, None


def lambda_loss(score_s: torch.Tensor, score_t: torch.Tensor, alpha: float = 1.0, eps: float = 1e-6) -> torch.Tensor:
    """criterion.py:148-189: teacher-ranked pairs weighted by |delta gain| * |delta discount| * 10 on
    softplus(-alpha * (s_i - s_j)); no (B, N, N) tensors."""
    return _GradLossFn.apply(score_s, score_t, "lambda", alpha, eps)


def ranknce_loss(score_s: torch.Tensor, score_t: torch.Tensor, temperature: float = 1.0,
                 lambda_weight: float = 1.0) -> torch.Tensor:
    """criterion.py:192-226: InfoNCE against the teacher's best page + lambda_weight * adjacent-pair ranking
    regulariser in teacher order; no sort, no gathers."""
    return _GradLossFn.apply(score_s, score_t, "ranknce", temperature, lambda_weight)


def infonce_supervised_loss(score_s: torch.Tensor, labels: torch.Tensor, temperature: float = 0.07) -> torch.Tensor:
    """criterion.py:43-54: a single library call (cross entropy) on the (B, N) scores."""
    return torch.nn.functional.cross_entropy(score_s / temperature, labels)


def infonce_distillation_loss(score_s: torch.Tensor, score_t: torch.Tensor, temperature: float = 0.07) -> torch.Tensor:
    """criterion.py:57-68: cross entropy against the teacher's arg-max page."""
    targets = torch.argmax(score_t.detach(), dim=1)
    return torch.nn.functional.cross_entropy(score_s / temperature, targets)


class RankGap(NamedTuple):
    gap: torch.Tensor          # (B, N) fp32  rank_t - rank_s
    G: torch.Tensor            # (N,)  fp32   sum_b |gap|
    rank_t: torch.Tensor       # (B, N) int32
    rank_s: torch.Tensor       # (B, N) int32
    topk_idx: torch.Tensor     # (B, kk) int64 teacher top-k, best first
    aux_doc_idx: torch.Tensor  # (B, a) int64 the a pages of the teacher top-k with the largest |gap|


def rank_gap(sc_t: torch.Tensor, sc_s: torch.Tensor, k: int = 10, aux_docs: int = 0) -> RankGap:
    """mainv3_iter_liscore_QA_hardtoken.py:368-385 in one launch (the reference runs four argsorts, two
    top-ks and two gathers).  kk = min(k, N), a = min(aux_docs, kk) like :376,380.  Equal scores keep
    index order (torch.argsort leaves their order unspecified)."""
    t, s = _rows(sc_t, "sc_t"), _rows(sc_s, "sc_s")
    if s.shape != t.shape:
        raise ValueError(f"sc_t {tuple(t.shape)} and sc_s {tuple(s.shape)} differ")
    B, n = t.shape
    dev = t.device
    kk = min(int(k), n)
    a = min(int(aux_docs), kk)
    gap = torch.empty((B, n), dtype=torch.float32, device=dev)
    G = torch.empty(n, dtype=torch.float32, device=dev)
    rt = torch.empty((B, n), dtype=torch.int32, device=dev)
    rs = torch.empty((B, n), dtype=torch.int32, device=dev)
    tk = torch.empty((B, kk), dtype=torch.int64, device=dev)
    aux = torch.empty((B, a), dtype=torch.int64, device=dev)
    lib = _lib.load()
    _lib.check(lib.evdr_rank_gap(_lib.handle(dev), t.data_ptr(), t.stride(0), s.data_ptr(), s.stride(0), B, n, kk, a,
                                 gap.data_ptr(), G.data_ptr(), rt.data_ptr(), rs.data_ptr(), tk.data_ptr(),
                                 aux.data_ptr() if a > 0 else None, _lib.stream_ptr(dev)), "evdr_rank_gap")
    return RankGap(gap, G, rt, rs, tk, aux)
