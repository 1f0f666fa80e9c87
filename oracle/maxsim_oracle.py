"""numpy restatement of the reference's MaxSim scorers (TEST INFRASTRUCTURE).

Follows, without sharing code with, the reference:

* ``maxsim_masked``      <- evaluator/retrieval.py:166-213 (score_multi_vector_masked)
* ``maxsim_list``        <- evaluator/retrieval.py:101-150 + left_padding :30-45
* ``l2_normalize``       <- utils/preprocess_data.py:8-9
* ``maxsim_backward``    <- the autograd of :166-213 (no source line)

All arithmetic is float32 like the reference (``Q.to(torch.float32)`` :176-177).
The page axis is processed one page at a time so that the (Q,C,Lq,Lp)
similarity tensor of the reference is never needed; results are independent of
the reference's ``chunk_p`` (verified by execution).
"""
from __future__ import annotations

import numpy as np

NEG_FILL = np.float32(-1e4)  # evaluator/retrieval.py:185


def _as_f32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def l2_normalize(x, eps: float = 1e-12) -> np.ndarray:
    """x / (||x||_2 + eps) along the last axis (utils/preprocess_data.py:8-9)."""
    x = _as_f32(x)
    nrm = np.sqrt((x * x).sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return (x / (nrm + np.float32(eps))).astype(np.float32)


def rowmax_argmax(Q, P, pmask):
    """Per (query, page, query token): masked max over page tokens and its first index.

    Returns ``(mx, am)`` with shapes (Nq, Np, Lq); ``mx`` holds -1e4 for pages with
    no valid token (as the reference's masked_fill does, :198), ``am`` the lowest
    index attaining the max (torch.max tie rule).
    """
    Q = _as_f32(Q)
    P = _as_f32(P)
    pmask = np.asarray(pmask).astype(bool)
    nq, lq, d = Q.shape
    npg, lp, _ = P.shape
    mx = np.empty((nq, npg, lq), dtype=np.float32)
    am = np.empty((nq, npg, lq), dtype=np.int64)
    Qf = Q.reshape(nq * lq, d)
    for c in range(npg):
        sim = Qf @ P[c].T  # (nq*lq, lp) float32
        sim = np.where(pmask[c][None, :], sim, NEG_FILL)
        a = sim.argmax(axis=1)
        mx[:, c, :] = sim[np.arange(sim.shape[0]), a].reshape(nq, lq)
        am[:, c, :] = a.reshape(nq, lq)
    return mx, am


def maxsim_masked(Q, P, qmask, pmask) -> np.ndarray:
    """out[q,c] = has[c] * sum_n qmask[q,n] * max_m (pmask[c,m] ? <Q[q,n],P[c,m]> : -1e4).

    evaluator/retrieval.py:166-213.  Masks of any dtype are taken as ``!= 0`` (:179-180).
    """
    qmask = np.asarray(qmask).astype(bool)
    pmask = np.asarray(pmask).astype(bool)
    mx, _ = rowmax_argmax(Q, P, pmask)
    has = pmask.any(axis=1).astype(np.float32)           # :192
    mx = mx * has[None, :, None]                          # :204
    mx = mx * qmask.astype(np.float32)[:, None, :]        # :207
    return mx.sum(axis=-1, dtype=np.float32).astype(np.float32)  # :210


def maxsim_backward(Q, P, qmask, pmask, gout):
    """Closed-form gradients of ``maxsim_masked`` w.r.t. P and Q.

    dP[c,m,:] = sum_{q,n} g[q,c] qmask[q,n] has[c] [m == argmax] Q[q,n,:]
    dQ[q,n,:] = sum_c     g[q,c] qmask[q,n] has[c] P[c, argmax, :]
    Ties go to the lowest index; masked tokens and empty pages receive exactly 0.
    """
    Q = _as_f32(Q)
    P = _as_f32(P)
    g = _as_f32(gout)
    qmask = np.asarray(qmask).astype(bool)
    pmask = np.asarray(pmask).astype(bool)
    nq, lq, d = Q.shape
    npg, lp, _ = P.shape
    _, am = rowmax_argmax(Q, P, pmask)
    has = pmask.any(axis=1)
    dP = np.zeros_like(P)
    dQ = np.zeros_like(Q)
    for c in range(npg):
        if not has[c]:
            continue
        for q in range(nq):
            w = g[q, c]
            for n in range(lq):
                if not qmask[q, n]:
                    continue
                m = am[q, c, n]
                dP[c, m] += w * Q[q, n]
                dQ[q, n] += w * P[c, m]
    return dP, dQ


def maxsim_list(qs, ps, batch_size: int = 128) -> np.ndarray:
    """List-input MaxSim without masks (evaluator/retrieval.py:101-150).

    Every block of ``batch_size`` queries / pages is zero-padded on the LEFT to the
    block's own longest sequence (:30-45) and the padded zero tokens take part in
    the max (a padded page's per-token max is max(true max, 0)); padded query
    tokens add max_m <0, p_m> = 0 for a page that also has padding, else the
    true max of a zero dot product, which is 0 as well.
    """
    if len(qs) == 0:
        raise ValueError("No queries provided")
    if len(ps) == 0:
        raise ValueError("No passages provided")
    qs = [_as_f32(q).reshape(-1, np.asarray(q).shape[-1]) for q in qs]
    ps = [_as_f32(p).reshape(-1, np.asarray(p).shape[-1]) for p in ps]
    out = np.zeros((len(qs), len(ps)), dtype=np.float32)
    for j0 in range(0, len(ps), batch_size):
        pblk = ps[j0:j0 + batch_size]
        lp = max(p.shape[0] for p in pblk)
        for jj, p in enumerate(pblk):
            padded = p.shape[0] < lp
            for i, q in enumerate(qs):
                sim = q @ p.T                                # (lq_i, lp_j)
                mx = sim.max(axis=1)
                if padded:
                    mx = np.maximum(mx, np.float32(0.0))      # left zero tokens of the page
                # left zero tokens of the query contribute max_m 0 = 0
                out[i, j0 + jj] = mx.sum(dtype=np.float32)
    return out


# ---------------------------------------------------------------------------
# torch-CPU variant used only as the timed CPU baseline (bench.py cpu_baseline /
# --impl reference).  Same arithmetic as the reference's hot loop -- one batched
# fp32 matmul per page chunk, fill, max, two multiplies, sum -- so that its
# throughput is representative of the reference on the same cores.
# ---------------------------------------------------------------------------
def maxsim_masked_torch_cpu(Q, P, qmask, pmask, chunk_p: int = 64):
    import torch

    Q = Q.to(torch.float32)
    P = P.to(torch.float32)
    qm = qmask.bool()
    pm = pmask.bool()
    nq, lq, d = Q.shape
    cols = []
    Qf = Q.reshape(nq * lq, d)
    qw = qm.to(torch.float32)
    for s in range(0, P.shape[0], chunk_p):
        Pc = P[s:s + chunk_p]
        mc = pm[s:s + chunk_p]
        c, lp, _ = Pc.shape
        sim = (Qf @ Pc.reshape(c * lp, d).T).reshape(nq, lq, c, lp)
        sim = torch.where(mc[None, None], sim, torch.tensor(-1e4, dtype=torch.float32))
        mx = sim.amax(dim=-1)                                    # (nq, lq, c)
        mx = mx * mc.any(dim=1).to(torch.float32)[None, None, :]
        mx = mx * qw[:, :, None]
        cols.append(mx.sum(dim=1))
    return torch.cat(cols, dim=1)
