"""numpy restatement of the (B, N) score-matrix consumers behind the scorer (TEST INFRASTRUCTURE).

Follows, without sharing code with, the reference:

* ``listwise_distillation_loss``  <- criterion.py:113-142
* ``score_preserving_loss``       <- criterion.py:73-83
* ``pairwise_distillation_loss``  <- criterion.py:89-110
* ``rank_gap``                    <- mainv3_iter_liscore_QA_hardtoken.py:368-385

Each loss also returns its gradient w.r.t. the student scores in closed form (the reference gets it
from autograd; ``tests/golden/losses.npz`` holds the autograd values of the unmodified reference
functions, against which this file is pinned in ``tests/test_oracle.py``).  float64 arithmetic: the
checker is more precise than both fp32 implementations it referees.
"""
from __future__ import annotations

import numpy as np


def _softmax(x):
    z = x - x.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def _desc_ranks(x):
    """argsort(argsort(x, descending)) with equal values in index order (stable)."""
    order = np.argsort(-x, axis=1, kind="stable")
    ranks = np.empty_like(order)
    rows = np.arange(x.shape[0])[:, None]
    ranks[rows, order] = np.arange(x.shape[1])[None, :]
    return ranks, order


def listwise_distillation_loss(score_s, score_t, k: int = 10, temperature: float = 1.0):
    """criterion.py:113-142.  Returns (loss, dloss/dscore_s)."""
    s = np.asarray(score_s, dtype=np.float64)
    t = np.asarray(score_t, dtype=np.float64)
    B, n = s.shape
    if k > n:
        raise ValueError("selected index k out of range")          # torch.topk, :132
    T = float(temperature)
    ps = _softmax(s / T)                                            # exp(log_softmax), :127
    pt = _softmax(t / T)                                            # :128
    _, order = _desc_ranks(t)                                       # top-k of prob_t == top-k of t, :132
    top = np.zeros((B, n), dtype=bool)
    top[np.arange(B)[:, None], order[:, :k]] = True
    sel_pt = np.where(top, pt, 0.0)                                 # :135
    loss = -(sel_pt * np.log(ps)).sum(axis=1).mean() * T * T        # :138-141
    grad = (T / B) * (ps * sel_pt.sum(axis=1, keepdims=True) - sel_pt)
    return loss, grad


def score_preserving_loss(score_s, score_t):
    """criterion.py:73-83 (F.mse_loss).  Returns (loss, dloss/dscore_s)."""
    s = np.asarray(score_s, dtype=np.float64)
    t = np.asarray(score_t, dtype=np.float64)
    d = s - t
    return (d * d).mean(), 2.0 * d / d.size


def pairwise_distillation_loss(score_s, score_t):
    """criterion.py:89-110.  Returns (loss, dloss/dscore_s)."""
    s = np.asarray(score_s, dtype=np.float64)
    t = np.asarray(score_t, dtype=np.float64)
    B, n = s.shape
    ds = s[:, :, None] - s[:, None, :]                              # :100
    dt = t[:, :, None] - t[:, None, :]                              # :101
    y = 1.0 / (1.0 + np.exp(-dt))                                   # :104
    bce = np.maximum(ds, 0.0) - ds * y + np.log1p(np.exp(-np.abs(ds)))   # :107
    sig = 1.0 / (1.0 + np.exp(-ds))
    grad = 2.0 * (sig - y).sum(axis=2) / (B * n * n)
    return bce.mean(), grad


def rank_gap(sc_t, sc_s, k: int = 10, aux_docs: int = 0):
    """mainv3_iter_liscore_QA_hardtoken.py:368-385.  Returns dict(gap, G, rank_t, rank_s, topk_idx,
    aux_doc_idx); equal values keep index order / teacher-rank order."""
    t = np.asarray(sc_t, dtype=np.float32)
    s = np.asarray(sc_s, dtype=np.float32)
    B, n = t.shape
    rank_t, order_t = _desc_ranks(t)                                # :369
    rank_s, _ = _desc_ranks(s)                                      # :370
    gap = rank_t.astype(np.float32) - rank_s.astype(np.float32)     # :371
    G = np.abs(gap).sum(axis=0)                                     # :372
    kk = min(int(k), n)                                             # :376
    topk_idx = order_t[:, :kk]                                      # :377
    gap_topk = np.abs(np.take_along_axis(gap, topk_idx, axis=1))    # :378
    a = min(int(aux_docs), kk)                                      # :379
    pos = np.argsort(-gap_topk, axis=1, kind="stable")[:, :a]       # :381
    aux = np.take_along_axis(topk_idx, pos, axis=1)                 # :382
    return {"gap": gap, "G": G, "rank_t": rank_t, "rank_s": rank_s, "topk_idx": topk_idx, "aux_doc_idx": aux}


def lambda_loss(score_s, score_t, alpha: float = 1.0, eps: float = 1e-6):
    """criterion.py:148-189.  Returns (loss, dloss/dscore_s)."""
    s = np.asarray(score_s, dtype=np.float64)
    t = np.asarray(score_t, dtype=np.float64)
    B, n = s.shape
    _, order = _desc_ranks(t)                                       # :162
    rows = np.arange(B)[:, None]
    ss, ts = s[rows, order], t[rows, order]                         # :163
    # The gain and discount factors are kept in float32 like the reference: MaxSim scores are ~10-30, where
    # sigmoid(t) = 1 - O(1e-6) and |sig(t_i) - sig(t_j)| is a small multiple of the fp32 spacing 6e-8 -- the
    # reference's value IS that quantised number, an exact sigmoid would differ from it by percents.
    f32 = np.float32
    disc = (f32(1.0) / np.log2(np.arange(1, n + 1, dtype=f32) + f32(1.0))).astype(f32)          # :167-168
    ddiff = np.abs(disc[None, :, None] - disc[None, None, :]).astype(np.float64)                # :171
    rel = (f32(1.0) / (f32(1.0) + np.exp(-ts.astype(f32)))).astype(f32)                         # :174
    w = np.abs(rel[:, :, None] - rel[:, None, :]).astype(np.float64) * ddiff * 10.0             # :175-178
    x = alpha * (ss[:, :, None] - ss[:, None, :])                   # :181
    sp = np.maximum(-x, 0.0) + np.log1p(np.exp(-np.abs(x)))         # -logsigmoid(x), :182
    mask = np.triu(np.ones((n, n)), k=1)                            # :186
    denom = mask.sum() + eps
    loss = (w * sp * mask).sum() / denom                            # :187
    dsp = -alpha / (1.0 + np.exp(x))                                # d softplus(-x) / d s_i
    gm = w * dsp * mask
    gs = (gm.sum(axis=2) - gm.sum(axis=1)) / denom                  # in sorted order
    grad = np.empty_like(gs)
    grad[rows, order] = gs
    return loss, grad


def ranknce_loss(score_s, score_t, temperature: float = 1.0, lambda_weight: float = 1.0):
    """criterion.py:192-226.  Returns (loss, dloss/dscore_s)."""
    s = np.asarray(score_s, dtype=np.float64)
    t = np.asarray(score_t, dtype=np.float64)
    B, n = s.shape
    _, order = _desc_ranks(t)                                       # :206
    rows = np.arange(B)[:, None]
    ss, ts = s[rows, order], t[rows, order]                         # :207
    p = _softmax(ss / temperature)
    infonce = -np.log(p[:, 0]).mean()                               # :210-211 (label 0)
    ds, dt = ss[:, :-1] - ss[:, 1:], ts[:, :-1] - ts[:, 1:]         # :214-215
    w = 1.0 / (1.0 + np.exp(-dt))                                   # :221
    sp = np.maximum(-ds, 0.0) + np.log1p(np.exp(-np.abs(ds)))       # softplus(-ds)
    ranking = (w * sp).mean() if n > 1 else 0.0                     # :222
    loss = infonce + lambda_weight * ranking                        # :224
    gs = p.copy()
    gs[:, 0] -= 1.0
    gs /= (B * temperature)
    if n > 1:
        d = -w / (1.0 + np.exp(ds)) * (lambda_weight / (B * (n - 1)))
        gs[:, :-1] += d
        gs[:, 1:] -= d
    grad = np.empty_like(gs)
    grad[rows, order] = gs
    return loss, grad
