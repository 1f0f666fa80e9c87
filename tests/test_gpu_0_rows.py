"""One short parity check per row of SURVEY.md section 8 (the coverage contract), collected FIRST.

The driver runs ``pytest -m gpu -x``: a single failure hides everything after it, so each row of the scope
table gets a sub-second check against the reference's golden vectors / the oracle here, before the broad
suites (test_gpu_a .. test_gpu_i) repeat them in depth.  Every check goes through the C ABI of libevdr.so.
"""
import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo
from oracle import maxsim_oracle as mo
from oracle import trec_oracle as to
from tests import _util as U

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded(evdr, cuda):
    """No silent fallback: the product path is libevdr.so (sm_100a) or an exception."""
    lib = evdr._lib.load()
    assert lib.evdr_version() >= 100
    assert torch.cuda.get_device_capability(cuda)[0] == 10, "this package targets sm_100a"


# --------------------------------------------------------------------------------------- A1
def test_A1_default_tcgen05_flavours_vs_reference_goldens(evdr, cuda):
    """score_multi_vector_masked, tcgen05 path: stored-input golden (9 x 12 x 200) and the seeded
    1030-token case, outputs of the unmodified reference."""
    z = U.load("medium")
    Q, P = U.bf16_from_i16(z["Q_bf16"]).to(cuda), U.bf16_from_i16(z["P_bf16"]).to(cuda)
    out = evdr.score_multi_vector_masked(Q, P, torch.from_numpy(z["qmask"]).to(cuda), torch.from_numpy(z["pmask"]).to(cuda), 64)
    U.assert_scores_close(out.cpu().numpy(), z["out"])
    Q, P, qm, pm, rel, zs = U.seeded_case("full1030")
    out = evdr.score_multi_vector_masked(Q.to(cuda), P.to(cuda), qm.to(cuda), pm.to(cuda))
    U.assert_scores_close(out.cpu().numpy(), zs["full1030/out"])
    assert (out[:, (~pm.any(1)).to(cuda)] == 0).all()
    # fp32 page tensors take the bf16x3 split by default: near-fp32 scores of the same golden
    assert evdr.retrieval.default_precision(P.float()) == "bf16x3" and evdr.retrieval.default_precision(P) == "bf16"
    out3 = evdr.score_multi_vector_masked(Q.float().to(cuda), P.float().to(cuda), qm.to(cuda), pm.to(cuda))
    np.testing.assert_allclose(out3.cpu().numpy(), zs["full1030/out"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)


@pytest.mark.parametrize("nq,mode", [(64, None), (100, None), (64, "pair"), (64, "1cta")])
def test_A1_bench_flavour_two_epilogue_groups_on_1040_slot_pages(evdr, cuda, nq, mode, monkeypatch):
    """The kernel flavour bench.py times (>= 16 query tiles, R = 2, two epilogue groups; uniform bf16 corpus of
    1040-slot pages used in place) against the numpy oracle on a sub-block and the exact fp32 kernel in full;
    plus the cta_group::2 and one-group flavours on the same data."""
    if mode:
        monkeypatch.setenv("EVDR_TC_MODE", mode)
    toks = evdr.synth.make_uniform_corpus(600, 1040, device=cuda, seed=17)
    pm = torch.ones(600, 1040, dtype=torch.bool, device=cuda)
    Q, qm, rel = evdr.synth.make_queries(toks, pm, nq, 32, seed=18, sigma=3.0, min_len=8, drop=0.1)
    pk = evdr.PackedPages.from_uniform_tokens(toks)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    got = pk.score_packed(Qb, qm32)
    flavour = evdr._lib.last_kernel()
    want = {None: "CG2=1,AM=0,KB=1,SP=0,EW=8" if nq % 16 == 0 else "CG2=0,AM=0,KB=1,SP=0,EW=8",      # 100 queries = 25 tiles
            "pair": "CG2=1,AM=0,KB=1,SP=0,EW=4", "1cta": "CG2=0,AM=0,KB=1,SP=0,EW=4"}[mode]
    assert want in flavour, flavour                              # the default IS the flavour bench.py times
    qs, cs = slice(3, 19), slice(101, 165)
    want = mo.maxsim_masked(Q[qs].float().cpu().numpy(), toks[cs].float().cpu().numpy(), qm[qs].cpu().numpy(),
                            pm[cs].cpu().numpy())
    U.assert_scores_close(got[qs, cs].cpu().numpy(), want)
    exact = evdr.score_multi_vector_masked(Q, toks, qm, pm, precision="fp32")
    U.assert_scores_close(got.cpu().numpy(), exact.cpu().numpy())
    # the drop-in call on the same tensors reaches the same kernel (zero-copy view of the corpus)
    evdr.retrieval.clear_pack_cache()
    assert torch.equal(evdr.score_multi_vector_masked(Q, toks, qm, pm, 64), got)


# --------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("name", ["comp64", "varlen", "lq1"])
def test_K2_masked_variable_length_compressed_pages(evdr, cuda, name):
    """Compressed 64-token pages, ragged lengths with holes, single-token queries: reference goldens through
    the packed stream (mask folded into the layout) and the short-page epilogue."""
    Q, P, qm, pm, rel, z = U.seeded_case(name)
    evdr.retrieval.clear_pack_cache()
    out = evdr.score_multi_vector_masked(Q.to(cuda), P.to(cuda), qm.to(cuda), pm.to(cuda), 64)
    U.assert_scores_close(out.cpu().numpy(), z[f"{name}/out"])
    pk = evdr.PackedPages.from_padded(P.to(cuda), pm.to(cuda), varlen=True)     # trimmed pages
    U.assert_scores_close(pk.score(Q.to(cuda), qm.to(cuda)).cpu().numpy(), z[f"{name}/out"])


# --------------------------------------------------------------------------------------- A2
def test_A2_list_scorer_vs_reference_golden(evdr, cuda):
    z = U.load("list_scorer")
    qs = [torch.from_numpy(a) for a in np.split(z["qs"], np.cumsum(z["q_lens"])[:-1])]
    ps = [torch.from_numpy(a) for a in np.split(z["ps"], np.cumsum(z["p_lens"])[:-1])]
    out = evdr.BaseVisualRetrieverProcessor.score_multi_vector(qs, ps, batch_size=3, device=str(cuda))
    assert out.device.type == "cpu" and out.dtype == torch.float32
    U.assert_scores_close(out.numpy(), z["out_b3"])
    with pytest.raises(ValueError, match="No queries provided"):
        evdr.BaseVisualRetrieverProcessor.score_multi_vector([], ps)


# --------------------------------------------------------------------------------------- A4 / f1
def test_A4_f1_masked_l2_normalize_and_fused_producer(evdr, cuda):
    g = torch.Generator().manual_seed(4)
    X = torch.randn(7, 41, 128, generator=g)
    m = torch.rand(7, 41, generator=g) > 0.2
    ref = mo.l2_normalize(X.numpy() * m.numpy()[..., None])
    Xa = X.to(cuda).requires_grad_(True)
    Y = evdr.l2_normalize_masked(Xa, m.to(cuda))
    np.testing.assert_allclose(Y.detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    Xb = X.clone().requires_grad_(True)
    xm = Xb * m.unsqueeze(-1)
    Yr = xm / (xm.norm(dim=-1, keepdim=True) + 1e-12)          # utils/preprocess_data.py:8-9 on P * mask
    gy = torch.randn(Y.shape, generator=g)
    Y.backward(gy.to(cuda))
    Yr.backward(gy)
    np.testing.assert_allclose(Xa.grad.cpu().numpy(), Xb.grad.numpy(), rtol=1e-4, atol=1e-5)
    # f1: the same normalisation fused into the bf16 packer
    pk = evdr.PackedPages.from_padded(X.to(cuda), m.to(cuda), normalize=True)
    got = pk.tokens.float().cpu().numpy().reshape(7, -1, 128)[:, :41]
    sel = m.numpy()
    np.testing.assert_allclose(got[sel], torch.from_numpy(ref[sel]).bfloat16().float().numpy(), rtol=0, atol=4e-3)


# --------------------------------------------------------------------------------------- A5
def test_A5_backward_vs_reference_autograd(evdr, cuda):
    z = U.load("medium")
    Q = U.bf16_from_i16(z["Q_bf16"]).float().to(cuda).requires_grad_(True)
    P = U.bf16_from_i16(z["P_bf16"]).float().to(cuda).requires_grad_(True)
    qm, pm = torch.from_numpy(z["qmask"]).to(cuda), torch.from_numpy(z["pmask"]).to(cuda)
    out = evdr.score_multi_vector_masked(Q, P, qm, pm, 5)
    np.testing.assert_allclose(out.detach().cpu().numpy(), z["out"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)
    out.backward(torch.from_numpy(z["gout"]).to(cuda))
    np.testing.assert_allclose(P.grad.cpu().numpy(), z["dP"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(Q.grad.cpu().numpy(), z["dQ"], rtol=1e-4, atol=1e-5)
    assert P.grad[~pm].abs().max() == 0, "masked page tokens receive exactly zero gradient"
    # exact fp32 route as well
    P2 = P.detach().clone().requires_grad_(True)
    evdr.score_multi_vector_masked(Q.detach(), P2, qm, pm, precision="fp32").backward(torch.from_numpy(z["gout"]).to(cuda))
    np.testing.assert_allclose(P2.grad.cpu().numpy(), z["dP"], rtol=1e-4, atol=1e-5)


# --------------------------------------------------------------------------------------- A6 / f4
def test_A6_f4_fused_losses_vs_reference_criterion(evdr, cuda):
    z = np.load(U.GOLDEN / "losses.npz")
    name = "b5_n70_k40_T2"
    B, N, k, aux, _ = (int(v) for v in z[f"{name}/spec"])
    T = float(z[f"{name}/T"])
    s, t = torch.from_numpy(z[f"{name}/s"]).to(cuda), torch.from_numpy(z[f"{name}/t"]).to(cuda)
    for key, fn in (("listwise", lambda x: evdr.listwise_distillation_loss(x, t, k=k, temperature=T)),
                    ("mse", lambda x: evdr.score_preserving_loss(x, t)),
                    ("pairwise", lambda x: evdr.pairwise_distillation_loss(x, t)),
                    ("ranknce", lambda x: evdr.ranknce_loss(x, t, temperature=T, lambda_weight=0.7)),
                    ("infonce_distill", lambda x: evdr.infonce_distillation_loss(x, t, temperature=0.07))):
        sg = s.clone().requires_grad_(True)
        loss = fn(sg)
        loss.backward()
        np.testing.assert_allclose(loss.item(), z[f"{name}/{key}_loss"], rtol=2e-5, err_msg=key)
        ref = z[f"{name}/{key}_grad"]
        np.testing.assert_allclose(sg.grad.cpu().numpy(), ref, rtol=2e-4, atol=2e-6 * np.abs(ref).max(), err_msg=key)
    sg = s.clone().requires_grad_(True)
    loss = evdr.lambda_loss(sg, t, alpha=0.5)
    np.testing.assert_allclose(loss.item(), z[f"{name}/lambda_loss"], rtol=1e-3)


# --------------------------------------------------------------------------------------- A7
def test_A7_topk_and_trec_measures(evdr, cuda):
    g = torch.Generator().manual_seed(3)
    s = torch.randn(6, 5000, generator=g)
    s[:, ::9] = s[:, 4:5]                                       # exact ties: column order decides
    vals, ids = evdr.top_k(s.to(cuda), 10)
    order = np.argsort(-s.numpy(), axis=1, kind="stable")[:, :10]
    assert np.array_equal(ids.cpu().numpy(), order)
    assert np.array_equal(vals.cpu().numpy(), np.take_along_axis(s.numpy(), order, 1))
    # evaluator (dict signature of evaluator/retrieval.py:230-255) vs the trec_eval / mteb restatement
    rng = np.random.default_rng(1)
    docs = [f"doc-{rng.integers(0, 10**6):06d}-{i}" for i in range(60)]
    sc = np.round(rng.standard_normal((20, 60)).astype(np.float32), 1)
    qrels = {f"q{q}": {docs[j]: int(rng.integers(1, 4)) for j in rng.choice(60, 2, replace=False)} for q in range(20)}
    qrels["judged-but-unscored"] = {docs[0]: 1}
    results = {f"q{q}": {docs[j]: float(sc[q, j]) for j in range(60)} for q in range(20)}
    ev = evdr.CustomRetrievalEvaluator()
    got, want = ev.compute_mteb_metrics(qrels, results), to.compute_mteb_metrics(qrels, results, ev.k_values)
    for fam in want:
        for key, v in want[fam].items():
            assert got[fam][key] == pytest.approx(v, abs=1.1e-5), (fam, key)
    assert got["NDCG"]["NDCG@5"] == want["NDCG"]["NDCG@5"] and got["Recall"]["Recall@1"] == want["Recall"]["Recall@1"]
    assert got["mRR"]["MRR@10"] == want["mRR"]["MRR@10"]      # divides by len(qrels) like mteb / BEIR


@pytest.mark.parametrize("nq", [1, 4, 5, 8, 13])
def test_A7_topk_fused_into_the_scan(evdr, cuda, nq):
    """north_star: 'a fused warp-level top-k'.  With <= 8 queries the scan keeps the k best pages per query in
    registers and only k candidates per CTA leave the SM (evdr_sharded_search): ids and scores bit-identical to
    evdr_topk / torch.topk over the full score row of the unfused scan (13 queries: the unfused route)."""
    toks = evdr.synth.make_uniform_corpus(3000, 1040, device=cuda, seed=31)
    toks[1500:1510] = toks[7:17]                               # duplicated pages: exact score ties, id order decides
    pm = torch.ones(3000, 1040, dtype=torch.bool, device=cuda)
    Q, qm, rel = evdr.synth.make_queries(toks, pm, nq, 32, seed=32, sigma=2.0, min_len=8, drop=0.1)
    pk = evdr.PackedPages.from_uniform_tokens(toks, page_offset=7000)
    full = pk.score(Q, qm)
    for k in (1, 10, 32):
        want_s, want_i = evdr.top_k(full, k, id_offset=7000)
        got_s, got_i = pk.search(Q, qm, k)
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s), (nq, k)
        ts, ti = torch.topk(full, k, dim=1)
        assert torch.equal(got_s, ts)
    if nq in (1, 13):                                          # k > 32: scores + two-stage selection inside the same call
        want_s, want_i = evdr.top_k(full, 50, id_offset=7000)
        got_s, got_i = pk.search(Q, qm, 50)
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
    # ragged pages, a corpus smaller than the grid, k beyond the page count
    P, pmask = evdr.synth.make_pages(37, 700, seed=5, device=cuda, n_holes=4, empty_frac=0.05, min_len=100)
    Qs, qms, _ = evdr.synth.make_queries(P, pmask, min(nq, 4), 32, seed=6, sigma=2.0)
    pks = evdr.PackedPages.from_padded(P, pmask, varlen=True)
    want_s, want_i = evdr.top_k(pks.score(Qs, qms), 32)
    got_s, got_i = pks.search(Qs, qms, 32)
    assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
    # compressed 64- and 16-token pages (several page ends per tile: the short-page epilogue feeds the list)
    for lp_c, n_c in ((64, 9000), (16, 20000)):
        Pc, pmc = evdr.synth.make_pages(n_c, lp_c, seed=lp_c, device=cuda, n_holes=2, empty_frac=0.01, min_len=lp_c // 2)
        Pc[4000:4005] = Pc[11:16]
        pmc[4000:4005] = pmc[11:16]                              # duplicated pages: ties by page id
        Qc, qmc, _ = evdr.synth.make_queries(Pc, pmc, min(nq, 4), 32, seed=lp_c + 1, sigma=1.5, min_len=8, drop=0.1)
        for varlen in (False, True):
            pkc = evdr.PackedPages.from_padded(Pc, pmc, varlen=varlen, page_offset=50)
            want_s, want_i = evdr.top_k(pkc.score(Qc, qmc), 10, id_offset=50)
            got_s, got_i = pkc.search(Qc, qmc, 10)
            assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s), (lp_c, varlen)
    small = evdr.PackedPages.from_padded(P[:5], pmask[:5], varlen=True)
    got_s, got_i = small.search(Qs, qms, 10)
    assert (got_i[:, 5:] == -1).all() and torch.isinf(got_s[:, 5:]).all()


def test_A7_graphed_search_replays(evdr, cuda):
    """pack -> scan + top-k -> merge captured once as a CUDA graph: every replay equals the eager search."""
    toks = evdr.synth.make_uniform_corpus(2500, 1040, device=cuda, seed=41)
    pm = torch.ones(2500, 1040, dtype=torch.bool, device=cuda)
    pk = evdr.PackedPages.from_uniform_tokens(toks)
    gs = evdr.GraphedSearch(evdr.ShardedSearch((0, 2500), mode="nccl"), pk, 1, 10)
    for seed in range(4):
        Q, qm, _ = evdr.synth.make_queries(toks, pm, 1, 32, seed=50 + seed, sigma=2.0, min_len=8)
        want_s, want_i = evdr.top_k(pk.score(Q, qm), 10)
        got_s, got_i = gs(Q.float(), qm)
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)


# --------------------------------------------------------------------------------------- A8
def test_A8_rank_gap_and_hard_token_mining(evdr, cuda):
    z = np.load(U.GOLDEN / "losses.npz")
    name = "b5_n70_k40_T2"
    B, N, k, aux, _ = (int(v) for v in z[f"{name}/spec"])
    s, t = torch.from_numpy(z[f"{name}/s"]).to(cuda), torch.from_numpy(z[f"{name}/t"]).to(cuda)
    rg = evdr.rank_gap(t, s, k=k, aux_docs=aux)
    for got, key in ((rg.rank_t, "rank_t"), (rg.rank_s, "rank_s"), (rg.gap, "gap"), (rg.G, "G"), (rg.topk_idx, "topk_idx")):
        np.testing.assert_array_equal(got.cpu().numpy(), z[f"{name}/{key}"], err_msg=key)
    # hard tokens: argmax_m max_n <q_n, p_m> over valid page tokens (mainv3_iter_liscore_QA_hardtoken.py:398-425)
    P, pm = evdr.synth.make_pages(20, 120, seed=3, device=cuda, n_holes=5, empty_frac=0.0, min_len=40)
    Q, qm, rel = evdr.synth.make_queries(P, pm, 8, 32, seed=4, sigma=2.0, min_len=6, drop=0.1)
    doc_idx = torch.randint(0, 20, (8, 4), generator=torch.Generator().manual_seed(5)).to(cuda)
    got = evdr.hard_token_argmax(Q, qm, P, pm, doc_idx).cpu().numpy()
    Qn, Pn, qmn, pmn = Q.float().cpu().numpy(), P.float().cpu().numpy(), qm.cpu().numpy(), pm.cpu().numpy()
    for qi in range(8):
        for a in range(4):
            c = int(doc_idx[qi, a])
            col = (Qn[qi][qmn[qi]] @ Pn[c].T).max(0)
            col[~pmn[c]] = -np.inf
            assert abs(col[got[qi, a]] - col.max()) < 1e-6


# --------------------------------------------------------------------------------------- f2 / f3
def test_f2_teacher_score_table(evdr, cuda):
    P, pm = evdr.synth.make_pages(50, 200, seed=21, device=cuda, n_holes=4, empty_frac=0.02, min_len=50)
    Q, qm, _ = evdr.synth.make_queries(P, pm, 70, 32, seed=22, sigma=2.0, min_len=6, drop=0.1)
    table = evdr.TeacherScoreTable(Q, qm, P, pm, query_block=32)
    idx = torch.tensor([5, 69, 0, 5, 33], device=cuda)
    want = mo.maxsim_masked(Q[idx].float().cpu().numpy(), P.float().cpu().numpy(), qm[idx].cpu().numpy(), pm.cpu().numpy())
    U.assert_scores_close(table[idx].cpu().numpy(), want)


def test_f3_object_arrays_to_device_layout(evdr, cuda):
    """utils/preprocess_data.py:37-105 (pad + valid & attn & img) folded into one ragged pack."""
    rng = np.random.default_rng(0)
    lens = [37, 1, 200, 64, 16]
    docs, attn, img = (np.empty(len(lens), dtype=object) for _ in range(3))
    for i, L in enumerate(lens):
        x = rng.standard_normal((L, 128)).astype(np.float32)
        docs[i] = x / np.linalg.norm(x, axis=-1, keepdims=True)
        attn[i] = rng.random(L) > 0.1
        img[i] = (rng.random(L) > 0.2).astype(np.int64)
    img[4] = np.zeros(16, dtype=bool)                           # a page without any valid token
    pk = evdr.PackedPages.from_object_arrays(docs, attn, img, device=cuda, k_blocks=3)
    Lmax = max(lens)
    P, pm = np.zeros((len(lens), Lmax, 128), np.float32), np.zeros((len(lens), Lmax), bool)
    for i, L in enumerate(lens):
        P[i, :L] = docs[i]
        pm[i, :L] = np.asarray(attn[i]).astype(bool) & np.asarray(img[i]).astype(bool)
    Q = torch.nn.functional.normalize(torch.randn(6, 20, 128, generator=torch.Generator().manual_seed(1)), dim=-1)
    qm = torch.ones(6, 20, dtype=torch.bool)
    want = mo.maxsim_masked(Q.numpy(), P, qm.numpy(), pm)
    got = pk.score(Q.to(cuda), qm.to(cuda)).cpu().numpy()
    U.assert_scores_close(got, want, rel=1e-5)
    assert (got[:, 4] == 0).all()


# --------------------------------------------------------------------------------------- (e)
def test_e_sharded_merge_on_one_rank_equals_global_topk(evdr, cuda):
    """Top-k of two half-corpora merged = global top-k (the exchange + merge step with world = 1 halves)."""
    g = torch.Generator().manual_seed(9)
    s = torch.randn(5, 4000, generator=g).to(cuda)
    v0, i0 = evdr.top_k(s[:, :2000].contiguous(), 10, id_offset=0)
    v1, i1 = evdr.top_k(s[:, 2000:].contiguous(), 10, id_offset=2000)
    mv, mi = evdr.topk_merge(torch.cat([v0, v1], 1), torch.cat([i0, i1], 1), 10)
    gv, gi = evdr.top_k(s, 10)
    assert torch.equal(mi, gi) and torch.equal(mv, gv)


def test_dispatch_picks_the_documented_flavour(evdr, cuda, monkeypatch):
    """which kernel runs for which shape (evdr_last_kernel)."""
    monkeypatch.delenv("EVDR_TC_MODE", raising=False)
    long_pages = evdr.PackedPages.from_uniform_tokens(evdr.synth.make_uniform_corpus(400, 1040, device=cuda, seed=3))
    short_pages = evdr.PackedPages.from_uniform_tokens(evdr.synth.make_uniform_corpus(4000, 64, device=cuda, seed=4))

    def flavour(pk, nq, search=False, k_blocks=1):
        Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=cuda), dim=-1)
        qm = torch.ones(nq, 32, dtype=torch.bool, device=cuda)
        pk.search(Q, qm, 10) if search else pk.score(Q, qm)
        return evdr._lib.last_kernel()

    assert "maxsim_scan_kernel<NQ=1>" in flavour(long_pages, 1)
    assert "maxsim_scan_kernel<NQ=2>" in flavour(long_pages, 2) and "topk=10" in flavour(long_pages, 2, search=True)
    assert "R=1,BN=256,STAGES=3,CG2=0" in flavour(long_pages, 4) and "TK=1" in flavour(long_pages, 4, search=True)
    assert "TA=1" in flavour(long_pages, 8) and "TA=1,TK=1" in flavour(long_pages, 8, search=True)
    assert "R=3,BN=256,STAGES=2,CG2=0" in flavour(long_pages, 12)
    assert "R=2,BN=256,STAGES=3,CG2=1" in flavour(long_pages, 16) and "EW=8" in flavour(long_pages, 16)
    assert "R=3,BN=256,STAGES=4,CG2=1" in flavour(long_pages, 24)              # six tiles: one pair group
    assert "CG2=0" in flavour(long_pages, 100)                                   # 25 tiles would idle 3 of 28 slots
    assert "CG2=1" in flavour(long_pages, 1024) and "TK=0" in flavour(long_pages, 1024, search=True)
    assert "SP=1,EW=4" in flavour(short_pages, 1) and "SP=1,EW=4,TA=0,TK=1" in flavour(short_pages, 3, search=True)
    assert "CG2=0" in flavour(short_pages, 8) and "SP=1,EW=8" in flavour(short_pages, 8)
    assert "CG2=1" in flavour(short_pages, 64) and "SP=1,EW=8" in flavour(short_pages, 64)
    x3 = evdr.PackedPages.from_padded(torch.randn(300, 200, 128, device=cuda), torch.ones(300, 200, dtype=torch.bool, device=cuda), k_blocks=3)
    assert "CG2=1,AM=0,KB=3" in flavour(x3, 32) and "BN=128,STAGES=2,CG2=0,AM=0,KB=3" in flavour(x3, 3)
