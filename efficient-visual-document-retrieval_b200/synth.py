"""Seeded synthetic ColPali-shape embeddings: 128-d, ~1030 tokens per page (or a
compressed 64), 32-token queries, all values bf16-representable so the reference's fp32 scorer and
the bf16 tensor-core path see identical inputs.  Relevance is planted so that nDCG@k is informative.
"""
from __future__ import annotations

import torch

BASE_SEED = 42  # the reference's default --seed (mainv2_iter_listwise.py:50)


def _unit_bf16(shape, gen, device):
    x = torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
    x = x / (x.norm(dim=-1, keepdim=True) + 1e-12)
    return x.to(torch.bfloat16)


def make_pages(n_pages: int, lp: int = 1030, d: int = 128, *, device="cpu", seed: int = BASE_SEED,
               n_holes: int = 6, empty_frac: float = 0.01, min_len: int | None = None,
               dtype=torch.bfloat16):
    """Pages (n_pages, lp, d) with zero right padding + bool mask.

    Mask = prefix-valid (length U[min_len, lp], default all lp) minus `n_holes` random "non-image"
    tokens, and `empty_frac` of the pages fully masked -- the arbitrary-pattern contract of
    utils/preprocess_data.py:101 (valid & attn & img)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    P = _unit_bf16((n_pages, lp, d), gen, device)
    if min_len is None:
        lens = torch.full((n_pages,), lp, device=device)
    else:
        lens = torch.randint(min_len, lp + 1, (n_pages,), generator=gen, device=device)
    ar = torch.arange(lp, device=device)
    mask = ar[None, :] < lens[:, None]
    if n_holes > 0:
        holes = torch.randint(0, lp, (n_pages, n_holes), generator=gen, device=device)
        mask.scatter_(1, holes, False)
    if empty_frac > 0:
        empty = torch.rand(n_pages, generator=gen, device=device) < empty_frac
        mask[empty] = False
    # zero right padding like pad_tokens_object (utils/preprocess_data.py:50-55)
    P = P * (ar[None, :] < lens[:, None])[..., None].to(P.dtype)
    return P.to(dtype), mask


def make_queries(P: torch.Tensor, pmask: torch.Tensor, n_queries: int, lq: int = 32, *,
                 sigma: float = 0.5, seed: int = BASE_SEED + 1, min_len: int | None = None,
                 drop: float = 0.0, dtype=torch.bfloat16):
    """Queries planted on random pages: token n of query q = normalize(P[rel(q), m_n] + sigma*eps).

    Returns (Q (nq, lq, d), qmask (nq, lq) bool, rel (nq,) int64)."""
    device = P.device
    gen = torch.Generator(device=device).manual_seed(seed)
    n_pages, lp, d = P.shape
    has = pmask.any(dim=1)
    cand = torch.nonzero(has).flatten()
    rel = cand[torch.randint(0, cand.numel(), (n_queries,), generator=gen, device=device)]
    # pick random VALID tokens of the relevant page
    w = pmask[rel].to(torch.float32) + 1e-9
    tok = torch.multinomial(w, lq, replacement=True, generator=gen)            # (nq, lq)
    base = P[rel[:, None], tok].to(torch.float32)                               # (nq, lq, d)
    noise = torch.randn(base.shape, generator=gen, device=device) / (d ** 0.5)
    Q = base + sigma * noise
    Q = Q / (Q.norm(dim=-1, keepdim=True) + 1e-12)
    if min_len is None:
        lens = torch.full((n_queries,), lq, device=device)
    else:
        lens = torch.randint(min_len, lq + 1, (n_queries,), generator=gen, device=device)
    qmask = torch.arange(lq, device=device)[None, :] < lens[:, None]
    if drop > 0:
        qmask &= torch.rand((n_queries, lq), generator=gen, device=device) >= drop
    Q = Q * (torch.arange(lq, device=device)[None, :] < lens[:, None])[..., None]
    return Q.to(torch.bfloat16).to(dtype), qmask, rel


def make_uniform_corpus(n_pages: int, tokens_per_page: int = 1040, *, device, seed: int = BASE_SEED,
                        chunk_pages: int = 4096) -> torch.Tensor:
    """All-valid bf16 corpus (n_pages, tokens_per_page, 128) generated on the device in chunks (never
    staged through host memory) -- the shard layout of the throughput configs."""
    out = torch.empty((n_pages, tokens_per_page, 128), dtype=torch.bfloat16, device=device)
    for i, c0 in enumerate(range(0, n_pages, chunk_pages)):
        c1 = min(n_pages, c0 + chunk_pages)
        gen = torch.Generator(device=device).manual_seed(seed + i)
        x = torch.randn((c1 - c0, tokens_per_page, 128), generator=gen, device=device, dtype=torch.float32)
        x = x / (x.norm(dim=-1, keepdim=True) + 1e-12)
        out[c0:c1] = x.to(torch.bfloat16)
        del x
    return out
