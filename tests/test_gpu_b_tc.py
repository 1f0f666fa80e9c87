"""GPU parity, part B: the tcgen05 / TMEM / TMA MaxSim kernel (bf16 operands, fp32 accumulation)
through the C ABI, against the reference's golden vectors, the oracle, and the exact GPU kernels.

Tolerance (north_star): scores within 1e-3 relative with fp32 accumulation -- written as
|got - ref| <= 1e-3 * max(|ref|, 1) in tests/_util.py; inputs are bf16-representable so both sides
see identical operands and the observed error is ~1e-6.
"""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as mo
from oracle import trec_oracle as to
from tests import _util as U

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("plain_bf16")]


@pytest.mark.parametrize("name", ["full1030", "comp64", "varlen", "student_mf5", "lq1"])
def test_seeded_tensor_core_path(evdr, cuda, name):
    Q, P, qm, pm, rel, z = U.seeded_case(name)
    evdr.retrieval.clear_pack_cache()
    out = evdr.score_multi_vector_masked(Q.to(cuda), P.to(cuda), qm.to(cuda), pm.to(cuda), 64)
    assert out.dtype == torch.float32 and out.shape == (Q.shape[0], P.shape[0])
    U.assert_scores_close(out.cpu().numpy(), z[f"{name}/out"])
    empty = ~pm.any(dim=1)
    assert (out[:, empty.to(cuda)] == 0).all(), "pages without valid tokens score exactly 0"
    # second call hits the packed-corpus cache and must give the same bits
    out2 = evdr.score_multi_vector_masked(Q.to(cuda), P.to(cuda), qm.to(cuda), pm.to(cuda))
    U.assert_scores_close(out2.cpu().numpy(), z[f"{name}/out"])


def test_medium_case(evdr, cuda):
    z = U.load("medium")
    Q = U.bf16_from_i16(z["Q_bf16"]).to(cuda)
    P = U.bf16_from_i16(z["P_bf16"]).to(cuda)
    out = evdr.score_multi_vector_masked(Q, P, torch.from_numpy(z["qmask"]).to(cuda),
                                         torch.from_numpy(z["pmask"]).to(cuda))
    U.assert_scores_close(out.cpu().numpy(), z["out"])


@pytest.mark.parametrize("mode", ["pair", "pair8", "1cta", "default", "sp8", "ew8"])
@pytest.mark.parametrize("nq", [1, 3, 4, 5, 8, 9, 12, 13, 16, 17, 24, 25, 32, 33, 40, 100])
def test_query_tile_configs_vs_exact_kernel(evdr, cuda, nq, mode, monkeypatch):
    """Every kernel flavour: one epilogue group with R = 1, 2, 3 resident query tiles ("1cta"), two
    epilogue groups with the tree fold ("ew8") or the end-mask / staged-sum epilogue ("sp8"), CTA pairs
    (cta_group::2) with R = 1..4 per CTA, and the default dispatch -- incl. partial last tiles, odd tile
    counts and partial last groups."""
    if mode == "default":
        monkeypatch.delenv("EVDR_TC_MODE", raising=False)
    else:
        monkeypatch.setenv("EVDR_TC_MODE", mode)
    P, pm = evdr.synth.make_pages(150, 1030, seed=100 + nq, device=cuda, n_holes=6, empty_frac=0.02)
    Q, qm, rel = evdr.synth.make_queries(P, pm, nq, 32, seed=200 + nq, sigma=3.0, min_len=8, drop=0.1)
    evdr.retrieval.clear_pack_cache()
    got = evdr.score_multi_vector_masked(Q, P, qm, pm)
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    U.assert_scores_close(got.cpu().numpy(), ref.cpu().numpy())
    assert got.abs().max() > 1.0


@pytest.mark.parametrize("lp,npg,varlen", [(64, 3000, False), (64, 3000, True), (333, 700, True), (16, 100, False),
                                            (1030, 333, True), (2100, 40, True), (1, 50, False)])
def test_page_shapes_vs_exact_kernel(evdr, cuda, lp, npg, varlen):
    P, pm = evdr.synth.make_pages(npg, lp, seed=lp, device=cuda, n_holes=min(3, lp - 1) if lp > 1 else 0,
                                  empty_frac=0.02, min_len=1)
    Q, qm, rel = evdr.synth.make_queries(P, pm, 21, 32, seed=lp + 1, sigma=2.0, min_len=4, drop=0.1)
    pk = evdr.PackedPages.from_padded(P, pm, varlen=varlen)
    got = pk.score(Q, qm)
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    U.assert_scores_close(got.cpu().numpy(), ref.cpu().numpy())


@pytest.mark.parametrize("lp,npg,varlen,nq", [(64, 3000, False, 64), (64, 2000, True, 128), (16, 5000, False, 48),
                                               (333, 700, True, 64), (208, 900, False, 16), (1, 300, False, 32)])
def test_short_pages_on_cta_pairs_vs_exact_kernel(evdr, cuda, lp, npg, varlen, nq, monkeypatch):
    """Query-tile counts that fill the four slots of a CTA pair on pages shorter than two tiles: cta_group::2 with
    the short-page (end-mask + staged-sum) epilogue -- the default there -- against the exact fp32 kernel, and
    bit-identical to the single-CTA short-page flavour."""
    P, pm = evdr.synth.make_pages(npg, lp, seed=lp + nq, device=cuda, n_holes=min(3, lp - 1) if lp > 1 else 0,
                                  empty_frac=0.02, min_len=1)
    Q, qm, rel = evdr.synth.make_queries(P, pm, nq, 32, seed=lp + 2, sigma=2.0, min_len=4, drop=0.1)
    pk = evdr.PackedPages.from_padded(P, pm, varlen=varlen)
    got = pk.score(Q, qm)
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    U.assert_scores_close(got.cpu().numpy(), ref.cpu().numpy())
    monkeypatch.setenv("EVDR_TC_MODE", "sp8")
    assert torch.equal(got, pk.score(Q, qm))


@pytest.mark.parametrize("ppi", [1, 3, 16, 1000])
def test_work_item_sizes(evdr, cuda, ppi):
    P, pm = evdr.synth.make_pages(203, 1030, seed=77, device=cuda)
    Q, qm, rel = evdr.synth.make_queries(P, pm, 30, 32, seed=78, sigma=3.0)
    pk = evdr.PackedPages.from_padded(P, pm)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    got = pk.score_packed(Qb, qm32, pages_per_item=ppi)
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    U.assert_scores_close(got.cpu().numpy(), ref.cpu().numpy())


def test_uniform_corpus_zero_copy_and_permutation(evdr, cuda):
    toks = evdr.synth.make_uniform_corpus(1200, 1040, device=cuda, seed=5)
    pm = torch.ones(1200, 1040, dtype=torch.bool, device=cuda)
    Q, qm, rel = evdr.synth.make_queries(toks, pm, 64, 32, seed=6, sigma=3.0)
    pk = evdr.PackedPages.from_uniform_tokens(toks)
    assert pk.tokens.data_ptr() == toks.data_ptr(), "uniform bf16 corpora are used in place"
    s = pk.score(Q, qm)
    s2 = evdr.PackedPages.from_padded(toks, pm).score(Q, qm)
    assert torch.equal(s, s2)
    # size-independent property: permuting pages permutes the scores
    perm = torch.randperm(1200, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    sp = evdr.PackedPages.from_uniform_tokens(toks[perm].contiguous()).score(Q, qm)
    assert torch.allclose(sp, s[:, perm], rtol=0, atol=1e-5)
    # subset against the exact kernel
    ref = evdr.score_multi_vector_masked(Q[:8], toks[:100], qm[:8], pm[:100], precision="fp32")
    U.assert_scores_close(s[:8, :100].cpu().numpy(), ref.cpu().numpy())
    # planted relevance is found
    top = s.argmax(dim=1)
    assert (top == rel).float().mean() > 0.5


def test_topk_and_ndcg_identical_to_reference_scores(evdr, cuda):
    """config 1 in miniature: top-k indices from our scores equal those from the reference's scores
    (apart from ties inside the tolerance) and nDCG@5 / Recall@1 are identical."""
    Q, P, qm, pm, rel, z = U.seeded_case("full1030")
    ref_scores = z["full1030/out"]
    got = evdr.score_multi_vector_masked(Q.to(cuda), P.to(cuda), qm.to(cuda), pm.to(cuda))
    k = 10
    _, ids = evdr.top_k(got, k)
    ids = ids.cpu().numpy()
    order = np.argsort(-ref_scores, axis=1, kind="stable")[:, :k]
    for q in range(ref_scores.shape[0]):
        for j in range(k):
            if ids[q, j] != order[q, j]:
                a, b = ref_scores[q, ids[q, j]], ref_scores[q, order[q, j]]
                assert abs(a - b) <= U.REL_TOL * max(abs(b), 1.0), "rank swap outside the tie tolerance"
    docs = [f"d{c:04d}" for c in range(P.shape[0])]
    qrels = {f"q{q}": {docs[int(rel[q])]: 1} for q in range(Q.shape[0])}
    res_ref = {f"q{q}": {docs[c]: float(ref_scores[q, c]) for c in range(P.shape[0])} for q in range(Q.shape[0])}
    gs = got.cpu().numpy()
    res_got = {f"q{q}": {docs[c]: float(gs[q, c]) for c in range(P.shape[0])} for q in range(Q.shape[0])}
    want = to.compute_mteb_metrics(qrels, res_ref)
    have = evdr.CustomRetrievalEvaluator().compute_mteb_metrics(qrels, res_got)
    assert have["NDCG"]["NDCG@5"] == want["NDCG"]["NDCG@5"]
    assert have["Recall"]["Recall@1"] == want["Recall"]["Recall@1"]
    assert 0.2 < want["NDCG"]["NDCG@5"] <= 1.0


def test_list_scorer_matches_reference(evdr, cuda):
    z = U.load("list_scorer")
    qs = [torch.from_numpy(a) for a in np.split(z["qs"], np.cumsum(z["q_lens"])[:-1])]
    ps = [torch.from_numpy(a) for a in np.split(z["ps"], np.cumsum(z["p_lens"])[:-1])]
    P = evdr.BaseVisualRetrieverProcessor
    out3 = P.score_multi_vector(qs, ps, batch_size=3, device=str(cuda))
    out128 = P.score_multi_vector(qs, ps, batch_size=128)
    assert out3.device.type == "cpu" and out3.dtype == torch.float32
    U.assert_scores_close(out3.numpy(), z["out_b3"])
    U.assert_scores_close(out128.numpy(), z["out_b128"])
    with pytest.raises(ValueError, match="No queries provided"):
        P.score_multi_vector([], ps)
    with pytest.raises(ValueError, match="No passages provided"):
        P.score_multi_vector(qs, [])


@pytest.mark.parametrize("where,dtype,d", [("cuda", torch.float32, 128), ("cpu", torch.bfloat16, 128),
                                            ("mixed", torch.float32, 128), ("cpu", torch.float32, 64)])
def test_list_scorer_input_forms(evdr, cuda, where, dtype, d):
    """Lists of GPU tensors, bf16 tensors, lists spread over devices and d != 128 (exact kernels) against the
    oracle's restatement of evaluator/retrieval.py:101-150 (itself pinned to the reference's outputs)."""
    from oracle import maxsim_oracle as mo
    g = torch.Generator().manual_seed(d)
    # bf16-representable values: both sides see identical inputs (the tensor-core path rounds to bf16)
    qs = [torch.nn.functional.normalize(torch.randn(int(n), d, generator=g), dim=-1).bfloat16().to(dtype)
          for n in torch.randint(3, 40, (9,), generator=g)]
    ps = [torch.nn.functional.normalize(torch.randn(int(n), d, generator=g), dim=-1).bfloat16().to(dtype)
          for n in torch.randint(1, 90, (23,), generator=g)]
    want = mo.maxsim_list([q.float().numpy() for q in qs], [p.float().numpy() for p in ps], batch_size=5)
    if where == "cuda":
        qs, ps = [q.to(cuda) for q in qs], [p.to(cuda) for p in ps]
    elif where == "mixed":
        qs = [q.to(cuda) if i % 2 else q for i, q in enumerate(qs)]
        ps = [p.to(cuda) if i % 3 else p for i, p in enumerate(ps)]
    got = evdr.BaseVisualRetrieverProcessor.score_multi_vector(qs, ps, batch_size=5, device=str(cuda))
    assert got.device.type == "cpu" and got.shape == (9, 23)
    U.assert_scores_close(got.numpy(), want)
    # a stacked (N, L, D) tensor is a valid "list" too (the reference iterates over it)
    st_q = torch.stack([q[:3] for q in qs if q.device == qs[1].device][:3])
    got2 = evdr.BaseVisualRetrieverProcessor.score_multi_vector(st_q, ps, batch_size=128, device=str(cuda))
    assert got2.shape == (st_q.shape[0], 23)


def test_search_api(evdr, cuda):
    P, pm = evdr.synth.make_pages(400, 64, seed=9, device=cuda, n_holes=0, empty_frac=0.0)
    Q, qm, rel = evdr.synth.make_queries(P, pm, 12, 32, seed=10, sigma=1.0)
    pk = evdr.PackedPages.from_padded(P, pm, page_offset=5000)
    s, ids = pk.search(Q, qm, k=10)
    assert (ids[:, 0] == rel + 5000).all()
    assert (s[:, :-1] >= s[:, 1:]).all()


@pytest.mark.parametrize("nq,ppc", [(1, 700), (3, 64), (21, 333)])
def test_host_corpus_search_equals_resident_search(evdr, cuda, nq, ppc):
    """Top-k over a corpus streamed from pinned host memory (no (nq, np) score matrix) = top-k of the resident scores."""
    toks = evdr.synth.make_uniform_corpus(2000, 208, device=cuda, seed=12)
    toks[1500:1504] = toks[3:7]                                   # score ties across chunks: page id decides
    pm = torch.ones(2000, 208, dtype=torch.bool, device=cuda)
    Q, qm, _ = evdr.synth.make_queries(toks, pm, nq, 32, seed=13, sigma=2.0, min_len=8, drop=0.1)
    hc = evdr.HostCorpus(toks.cpu().pin_memory(), cuda, pages_per_chunk=ppc, page_offset=1000)
    want_s, want_i = evdr.top_k(evdr.PackedPages.from_uniform_tokens(toks).score(Q, qm), 10, id_offset=1000)
    got_s, got_i = hc.search(Q, qm, 10)
    assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)


@pytest.mark.parametrize("lq", [33, 50, 64, 100, 128])
def test_long_queries_split_into_groups(evdr, cuda, lq):
    """Queries of up to 128 tokens stay on the tensor-core path as 32-token groups (exact split)."""
    P, pm = evdr.synth.make_pages(120, 333, seed=lq, device=cuda, n_holes=4, empty_frac=0.02, min_len=50)
    Q, qm, rel = evdr.synth.make_queries(P, pm, 11, lq, seed=lq + 1, sigma=2.0, min_len=lq // 2, drop=0.1)
    got = evdr.score_multi_vector_masked(Q, P, qm, pm)
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    U.assert_scores_close(got.cpu().numpy(), ref.cpu().numpy())
    assert got.shape == (11, 120)


@pytest.mark.parametrize("nq,ppc", [(5, 7), (40, 16), (3, 1000)])
def test_host_corpus_streaming_equals_resident_scoring(evdr, cuda, nq, ppc):
    """HostCorpus streams a pinned host corpus through two staging buffers while the kernel scores the
    previous chunk: per-page results are bit-identical to scoring the resident corpus in one launch."""
    npg, L = 53, 208
    corpus = evdr.synth.make_uniform_corpus(npg, L, device=cuda, seed=9)
    host = corpus.cpu().pin_memory()
    g = torch.Generator(device=cuda).manual_seed(nq)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g, device=cuda), dim=-1)
    qm = torch.rand(nq, 32, generator=g, device=cuda) > 0.1
    want = evdr.PackedPages.from_uniform_tokens(corpus).score(Q, qm)
    hc = evdr.HostCorpus(host, cuda, pages_per_chunk=ppc)
    got = hc.score(Q, qm)
    assert torch.equal(got, want)
    out = torch.empty_like(want)
    assert hc.score(Q, qm, out=out) is out and torch.equal(out, want)      # staging buffers reused
    with pytest.raises(ValueError):
        evdr.HostCorpus(corpus.cpu(), cuda)                                   # not pinned


@pytest.mark.parametrize("nq,lp,npg", [(61, 1030, 150), (64, 1030, 97), (100, 1030, 150), (131, 1030, 60),
                                       (5, 64, 3000), (40, 64, 3000), (100, 333, 700), (9, 16, 500), (64, 208, 900)])
def test_tmem_resident_query_tiles_vs_exact_kernel(evdr, cuda, nq, lp, npg, monkeypatch):
    """EVDR_TC_MODE=ta: the query tiles (A operand) are staged into TENSOR MEMORY by the epilogue warps and the
    MMAs read A from TMEM (192-token page tiles, 4-stage ring) -- long pages with the tree / segment fold and
    short pages with the end-mask epilogue, partial last tiles and groups included."""
    monkeypatch.setenv("EVDR_TC_MODE", "ta")
    P, pm = evdr.synth.make_pages(npg, lp, seed=300 + nq, device=cuda, n_holes=min(4, lp - 1), empty_frac=0.02,
                                  min_len=max(1, lp // 3))
    Q, qm, rel = evdr.synth.make_queries(P, pm, nq, 32, seed=400 + nq, sigma=3.0, min_len=8, drop=0.1)
    evdr.retrieval.clear_pack_cache()
    got = evdr.PackedPages.from_padded(P, pm, varlen=True).score(Q, qm)
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    U.assert_scores_close(got.cpu().numpy(), ref.cpu().numpy())
    assert got.abs().max() > 1.0
