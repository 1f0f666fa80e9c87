"""GPU parity at BASELINE.json's FULL sizes, through properties that do not need a full-size oracle run:
random sub-blocks against the exact fp32 kernel and the numpy oracle, duplicated pages, page-order
permutation, sharding (local top-k + merge == global top-k) and top-k against torch.topk on the same
scores.  One B200 (180 GB) holds every case; the largest corpus is 26.6 GB."""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as mo
from tests import _util as U

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("plain_bf16")]


def _need(cuda, gib):
    import gc
    gc.collect()
    torch.cuda.empty_cache()                                   # blocks cached by earlier tests count as used
    free, _ = torch.cuda.mem_get_info(cuda)
    if free < gib * (1 << 30):
        pytest.skip(f"needs {gib} GiB of free HBM")


def _queries(nq, cuda, seed, drop=0.1):
    g = torch.Generator(device=cuda).manual_seed(seed)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g, device=cuda), dim=-1)
    Q = Q.bfloat16().float()                                   # bf16-representable: both sides see the same inputs
    qm = torch.rand(nq, 32, generator=g, device=cuda) >= drop
    return Q, qm


def _spot_check(evdr, scores, corpus, Q, qm, n_pages, seed):
    """Random pages x all queries against the exact fp32 kernel; a corner of it against the numpy oracle."""
    g = torch.Generator(device=corpus.device).manual_seed(seed)
    idx = torch.randperm(corpus.shape[0], generator=g, device=corpus.device)[:n_pages]
    sub = corpus[idx].contiguous()
    pm = torch.ones(sub.shape[:2], dtype=torch.bool, device=corpus.device)
    ref = evdr.score_multi_vector_masked(Q, sub, qm, pm, precision="fp32")
    U.assert_scores_close(scores[:, idx].cpu().numpy(), ref.cpu().numpy())
    o = mo.maxsim_masked(Q[:6].cpu().numpy(), sub[:5].float().cpu().numpy(), qm[:6].cpu().numpy(), pm[:5].cpu().numpy())
    U.assert_scores_close(scores[:6, idx[:5]].cpu().numpy(), o)


def test_config3_1024_queries_x_100k_pages(evdr, cuda):
    """BASELINE configs[2] at full size: the bench workload itself."""
    _need(cuda, 40)
    npg = 100_000
    corpus = evdr.synth.make_uniform_corpus(npg, 1040, device=cuda, seed=42)
    corpus[:, 1030:] = corpus[:, 1029:1030]                    # 1030 valid tokens, filler slots duplicate one
    corpus[77_777] = corpus[123]                               # a duplicated page far away
    Q, qm = _queries(1024, cuda, 43)
    pk = evdr.PackedPages.from_uniform_tokens(corpus)
    scores = pk.score(Q, qm)
    assert scores.shape == (1024, npg) and bool(torch.isfinite(scores).all())
    assert torch.equal(scores[:, 77_777], scores[:, 123]), "identical pages score identically wherever they sit"
    _spot_check(evdr, scores, corpus, Q, qm, 96, seed=1)
    # masked query tokens contribute nothing: an all-masked query scores exactly 0 everywhere
    qm0 = qm.clone()
    qm0[5] = False
    s5 = pk.score(Q[4:8], qm0[4:8])
    assert (s5[1] == 0).all() and torch.allclose(s5[0], scores[4], rtol=0, atol=2e-5)
    # top-10 against torch.topk on the very same scores (ties inside the tolerance may swap ids)
    top_s, top_i = evdr.top_k(scores, 10)
    ts, ti = torch.topk(scores, 10, dim=1)
    assert torch.equal(top_s, ts)
    same = (top_i == ti)
    assert same.float().mean() >= 0.999
    assert torch.equal(scores.gather(1, top_i), ts), "any differing id is an exact score tie"
    # sharding property: top-10 of two half-corpora merged == global top-10
    h = npg // 2
    a_s, a_i = evdr.top_k(scores[:, :h], 10)
    b_s, b_i = evdr.top_k(scores[:, h:], 10, id_offset=h)
    m_s, m_i = evdr.topk_merge(torch.cat([a_s, b_s], 1), torch.cat([a_i, b_i], 1), 10)
    assert torch.equal(m_s, top_s) and torch.equal(m_i, top_i)
    # the reference-facing call on the same tensors (zero-copy packed view of the bf16 corpus)
    pm = torch.ones((npg, 1040), dtype=torch.bool, device=cuda)
    s2 = evdr.score_multi_vector_masked(Q[:64], corpus, qm[:64], pm, 64)
    assert torch.allclose(s2, scores[:64], rtol=0, atol=2e-5)


def test_config4_single_query_scan_200k_pages(evdr, cuda):
    """BASELINE configs[3], the 1-GPU point of the sharded 1M-page corpus scaled to what one test may hold:
    a single-query streaming scan + top-10 over 200k pages, page order permuted -> same ranking."""
    _need(cuda, 120)
    npg = 200_000
    corpus = evdr.synth.make_uniform_corpus(npg, 1040, device=cuda, seed=7)
    Q, qm = _queries(1, cuda, 8, drop=0.0)
    pk = evdr.PackedPages.from_uniform_tokens(corpus)
    s = pk.score(Q, qm)
    _spot_check(evdr, s, corpus, Q, qm, 64, seed=2)
    top_s, top_i = pk.search(Q, qm, 10)
    ts, ti = torch.topk(s, 10, dim=1)
    assert torch.equal(top_s, ts) and torch.equal(top_i, ti)
    # reversed page order: the same pages win with the same scores
    rev = torch.flip(corpus, dims=[0])
    rs, ri = evdr.PackedPages.from_uniform_tokens(rev).search(Q, qm, 10)
    assert torch.allclose(rs, top_s, rtol=0, atol=1e-5)
    assert torch.equal(npg - 1 - ri, top_i)
    # 64-query batch on the same corpus: row 0 of the batch equals the single-query scan
    Q64, qm64 = _queries(64, cuda, 9)
    Q64[0], qm64[0] = Q[0], qm[0]
    s64 = pk.score(Q64, qm64)
    assert torch.allclose(s64[0], s[0], rtol=0, atol=2e-5)


def test_config2_compressed_pages_at_scale(evdr, cuda):
    """Compressed 64-token pages (BASELINE configs[1]) on a corpus far beyond L2: 400 k pages x 64 token
    slots (6.5 GB) with ragged valid lengths, holes and empty pages through the masked variable-length
    packer."""
    _need(cuda, 60)
    npg, lp = 400_000, 64
    g = torch.Generator(device=cuda).manual_seed(3)
    P = torch.nn.functional.normalize(torch.randn(npg, lp, 128, generator=g, device=cuda), dim=-1).bfloat16()
    lens = torch.randint(1, lp + 1, (npg,), generator=g, device=cuda)
    pm = torch.arange(lp, device=cuda)[None, :] < lens[:, None]
    pm[torch.rand(npg, generator=g, device=cuda) < 0.01] = False          # 1 % empty pages
    pm &= torch.rand(npg, lp, generator=g, device=cuda) > 0.05             # holes: arbitrary mask pattern
    Q, qm = _queries(256, cuda, 4)
    pk = evdr.PackedPages.from_padded(P, pm, varlen=True)
    s = pk.score(Q, qm)
    empty = ~pm.any(dim=1)
    assert (s[:, empty] == 0).all(), "pages without valid tokens score exactly 0 (:192,204)"
    idx = torch.randperm(npg, generator=g, device=cuda)[:300]
    ref = evdr.score_multi_vector_masked(Q, P[idx], qm, pm[idx], precision="fp32")
    U.assert_scores_close(s[:, idx].cpu().numpy(), ref.cpu().numpy())
    o = mo.maxsim_masked(Q[:4].cpu().numpy(), P[idx[:8]].float().cpu().numpy(), qm[:4].cpu().numpy(),
                         pm[idx[:8]].cpu().numpy())
    U.assert_scores_close(s[:4, idx[:8]].cpu().numpy(), o)
    # rectangular plan (no trimming) gives the same scores
    s_rect = evdr.PackedPages.from_padded(P, pm, varlen=False).score(Q[:32], qm[:32])
    assert torch.allclose(s_rect, s[:32], rtol=0, atol=2e-5)
