"""GPU parity, part A: exact fp32 kernels, backward, packing, top-k, metrics -- all through the
C ABI of libevdr.so (ctypes), checked against the oracle and the reference's golden vectors."""
import importlib
import math
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as mo
from oracle import trec_oracle as to
from tests import _util as U

pytestmark = pytest.mark.gpu


def _dev(t, cuda):
    return t.to(cuda)


# ----------------------------------------------------------------------------- forward, exact path
def test_micro_cases(evdr, cuda):
    z = U.load("micro")
    Q, P = torch.from_numpy(z["Q"]).to(cuda), torch.from_numpy(z["P"]).to(cuda)
    qm, pm = torch.from_numpy(z["qmask"]).to(cuda), torch.from_numpy(z["pmask"]).to(cuda)
    out = evdr.score_multi_vector_masked(Q, P, qm, pm, 2)          # d=8 -> exact kernels
    assert out.dtype == torch.float32 and out.device == Q.device and out.shape == (3, 4)
    np.testing.assert_allclose(out.cpu().numpy(), z["out"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)
    assert (out[:, 2] == 0).all() and (out[1] == 0).all()
    # bf16 inputs, integer masks, keyword chunk_p, 4-argument form
    o2 = evdr.score_multi_vector_masked(Q.bfloat16(), P.bfloat16(), qm, pm, chunk_p=7)
    np.testing.assert_allclose(o2.cpu().numpy(), z["out_bf16_in"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)
    o3 = evdr.score_multi_vector_masked(Q, P, qm.long(), pm.int())
    np.testing.assert_allclose(o3.cpu().numpy(), z["out_int_masks"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)


def test_micro_backward(evdr, cuda):
    z = U.load("micro")
    Q = torch.from_numpy(z["Q"]).to(cuda).requires_grad_(True)
    P = torch.from_numpy(z["P"]).to(cuda).requires_grad_(True)
    qm, pm = torch.from_numpy(z["qmask"]).to(cuda), torch.from_numpy(z["pmask"]).to(cuda)
    out = evdr.score_multi_vector_masked(Q, P, qm, pm, 3)
    assert out.requires_grad
    out.backward(torch.from_numpy(z["gout"]).to(cuda))
    np.testing.assert_allclose(P.grad.cpu().numpy(), z["dP"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(Q.grad.cpu().numpy(), z["dQ"], rtol=1e-5, atol=1e-6)
    assert P.grad[~pm].abs().max() == 0, "masked page tokens must receive exactly zero gradient"


def test_medium_forward_backward(evdr, cuda):
    z = U.load("medium")
    Q = U.bf16_from_i16(z["Q_bf16"]).float().to(cuda).requires_grad_(True)
    P = U.bf16_from_i16(z["P_bf16"]).float().to(cuda).requires_grad_(True)
    qm, pm = torch.from_numpy(z["qmask"]).to(cuda), torch.from_numpy(z["pmask"]).to(cuda)
    out = evdr.score_multi_vector_masked(Q, P, qm, pm, 5)
    np.testing.assert_allclose(out.detach().cpu().numpy(), z["out"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)
    out.backward(torch.from_numpy(z["gout"]).to(cuda))
    np.testing.assert_allclose(P.grad.cpu().numpy(), z["dP"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(Q.grad.cpu().numpy(), z["dQ"], rtol=1e-4, atol=1e-5)
    # determinism: no floating-point atomics anywhere in the backward
    P2 = P.detach().clone().requires_grad_(True)
    evdr.score_multi_vector_masked(Q.detach(), P2, qm, pm).backward(torch.from_numpy(z["gout"]).to(cuda))
    assert torch.equal(P2.grad, P.grad)


@pytest.mark.parametrize("name", U.SEEDED_CASES)
def test_seeded_exact(evdr, cuda, name):
    Q, P, qm, pm, rel, z = U.seeded_case(name)
    before = (Q.clone(), P.clone())
    Qd, Pd = Q.to(cuda), P.to(cuda)
    out = evdr.score_multi_vector_masked(Qd, Pd, qm.to(cuda), pm.to(cuda), 64, precision="fp32")
    np.testing.assert_allclose(out.cpu().numpy(), z[f"{name}/out"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)
    assert torch.equal(Qd.cpu(), before[0]) and torch.equal(Pd.cpu(), before[1]), "inputs must not be mutated"


def test_student_backward(evdr, cuda):
    Q, P, qm, pm, rel, z = U.seeded_case("student_mf5")
    Pg = P.float().to(cuda).requires_grad_(True)
    out = evdr.score_multi_vector_masked(Q.to(cuda), Pg, qm.to(cuda), pm.to(cuda), 32)
    np.testing.assert_allclose(out.detach().cpu().numpy(), z["student_mf5/out"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)
    out.backward(torch.from_numpy(z["student_mf5/gout"]).to(cuda))
    np.testing.assert_allclose(Pg.grad.cpu().numpy(), z["student_mf5/dP"], rtol=1e-4, atol=1e-5)
    assert Pg.grad[~pm.to(cuda)].abs().max() == 0


def test_no_grad_and_odd_shapes(evdr, cuda):
    g = torch.Generator().manual_seed(5)
    Q = torch.randn(5, 37, 20, generator=g)           # lq > 32, d not 128
    P = torch.randn(11, 70, 20, generator=g)
    qm = torch.rand(5, 37, generator=g) > 0.3
    pm = torch.rand(11, 70, generator=g) > 0.3
    with torch.no_grad():
        out = evdr.score_multi_vector_masked(Q.to(cuda), P.to(cuda), qm.to(cuda), pm.to(cuda))
    ref = mo.maxsim_masked(Q.numpy(), P.numpy(), qm.numpy(), pm.numpy())
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)
    assert not out.requires_grad


def test_cpu_inputs_raise(evdr):
    Q = torch.randn(2, 4, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        evdr.score_multi_vector_masked(Q, Q, torch.ones(2, 4), torch.ones(2, 4))


# ----------------------------------------------------------------------------- fused l2 normalise
@pytest.mark.parametrize("n,l,d", [(6, 50, 128), (3, 7, 128), (2, 4, 128), (5, 9, 64), (1000, 41, 128)])
def test_l2_normalize_masked(evdr, cuda, n, l, d):
    """d == 128 runs the register-resident 4-tokens-per-warp kernels (incl. token counts that are not a
    multiple of 4), other dims the generic one-warp-per-token kernels."""
    g = torch.Generator().manual_seed(9)
    X = torch.randn(n, l, d, generator=g).to(cuda)
    X[1, 3] = 0
    m = (torch.rand(n, l, generator=g) > 0.2).to(cuda)
    Xa = X.clone().requires_grad_(True)
    Xb = X.clone().requires_grad_(True)
    Y = evdr.l2_normalize_masked(Xa, m)
    xm = Xb * m.unsqueeze(-1)
    Yr = xm / (xm.norm(dim=-1, keepdim=True) + 1e-12)   # utils/preprocess_data.py:8-9 on P*mask
    np.testing.assert_allclose(Y.detach().cpu().numpy(), Yr.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    gy = torch.randn(Y.shape, generator=g).to(cuda)
    Y.backward(gy)
    Yr.backward(gy)
    sel = torch.ones(n, l, dtype=torch.bool, device=cuda)
    sel[1, 3] = False  # zero vector: torch's norm backward is 0/0-guarded differently
    np.testing.assert_allclose(Xa.grad[sel].cpu().numpy(), Xb.grad[sel].cpu().numpy(), rtol=1e-4, atol=1e-5)


# ----------------------------------------------------------------------------- packed page stream
@pytest.mark.parametrize("varlen", [False, True])
def test_pack_pages_invariants(evdr, cuda, varlen):
    P, pm = evdr.synth.make_pages(37, 333, seed=3, n_holes=9, empty_frac=0.1, min_len=1)
    pk = evdr.PackedPages.from_padded(P.to(cuda), pm.to(cuda), varlen=varlen)
    off = pk.chunk_off.cpu().numpy()
    toks = pk.tokens.float().cpu().numpy().reshape(-1, 128)
    assert off[0] == 0 and off[-1] == pk.total_chunks and toks.shape[0] == pk.total_chunks * 16
    Pn, pmn = P.float().numpy(), pm.numpy()
    for c in range(37):
        nslot = (off[c + 1] - off[c]) * 16
        valid = np.nonzero(pmn[c])[0]
        if varlen:
            want = max(1, math.ceil((valid.max() + 1) / 16)) if valid.size else 1
        else:
            want = math.ceil(333 / 16)
        assert off[c + 1] - off[c] == want
        page = toks[off[c] * 16: off[c] * 16 + nslot]
        if valid.size == 0:
            assert (page == 0).all(), "page without valid tokens must be zero tokens"
            continue
        vt = Pn[c][valid]
        for m in range(nslot):
            if m < 333 and pmn[c, m]:
                assert (page[m] == Pn[c, m]).all()
            else:   # filler: copy of the nearest preceding valid token, else the first valid
                prev = valid[valid <= m]
                src = prev.max() if prev.size else valid.min()
                assert (page[m] == Pn[c, src]).all()


def test_pack_normalize_fuses_l2(evdr, cuda):
    g = torch.Generator().manual_seed(4)
    X = torch.randn(5, 64, 128, generator=g)
    m = torch.rand(5, 64, generator=g) > 0.2
    pk = evdr.PackedPages.from_padded(X.to(cuda), m.to(cuda), normalize=True)
    ref = mo.l2_normalize(X.numpy() * m.numpy()[..., None])
    got = pk.tokens.float().cpu().numpy().reshape(5, 64, 128)
    sel = m.numpy()
    np.testing.assert_allclose(got[sel], torch.from_numpy(ref[sel]).bfloat16().float().numpy(), rtol=0, atol=4e-3)


# ----------------------------------------------------------------------------- top-k
def _oracle_topk(scores, k, tie=None):
    nq, n = scores.shape
    tie = np.arange(n) if tie is None else tie
    ids = np.full((nq, k), -1, dtype=np.int64)
    vals = np.full((nq, k), -np.inf, dtype=np.float32)
    for q in range(nq):
        order = sorted(range(n), key=lambda c: (-scores[q, c], tie[c]))[:k]
        ids[q, :len(order)] = order
        vals[q, :len(order)] = scores[q, order]
    return vals, ids


@pytest.mark.parametrize("nq,n,k", [(7, 500, 10), (3, 70, 100), (1, 200000, 10), (64, 3000, 128), (2, 5, 5),
                                    (5, 1000, 1), (4, 4097, 32), (3, 31, 32), (2, 129, 17), (1, 125000, 10),
                                    (70, 100000, 10), (300, 1234, 33)])
def test_topk_matches_sort(evdr, cuda, nq, n, k):
    g = torch.Generator().manual_seed(nq * 1000 + k)
    s = torch.randn(nq, n, generator=g)
    s[:, ::7] = s[:, 3:4]                    # plenty of exact ties
    vals, ids = evdr.top_k(s.to(cuda), k)
    rv, ri = _oracle_topk(s.numpy(), k)
    assert np.array_equal(ids.cpu().numpy(), ri)
    assert np.array_equal(vals.cpu().numpy(), rv)


def test_topk_minus_inf_scores_and_register_list_edge_cases(evdr, cuda):
    """k <= 32 takes the register-list kernel: -inf scores are real candidates (absent pages of a result
    dict are -inf, evaluator path), all-equal rows resolve by column, descending rows insert at the tail and
    ascending rows at the head every time."""
    n = 700
    rows = [torch.full((n,), float("-inf")), torch.zeros(n), torch.arange(n, 0, -1).float(),
            torch.arange(n).float(), torch.cat([torch.full((n - 3,), float("-inf")), torch.tensor([1.0, 3.0, 2.0])])]
    s = torch.stack(rows)
    for k in (1, 10, 32):
        vals, ids = evdr.top_k(s.to(cuda), k)
        rv, ri = _oracle_topk(s.numpy(), k)
        assert np.array_equal(ids.cpu().numpy(), ri), k
        assert np.array_equal(vals.cpu().numpy(), rv), k


def test_topk_tie_rank_and_offset(evdr, cuda):
    g = torch.Generator().manual_seed(1)
    s = torch.randint(0, 4, (9, 300), generator=g).float()     # massive ties
    tie = torch.randperm(300, generator=g).int()
    vals, ids = evdr.top_k(s.to(cuda), 20, tie_rank=tie.to(cuda), id_offset=1000)
    rv, ri = _oracle_topk(s.numpy(), 20, tie.numpy())
    assert np.array_equal(ids.cpu().numpy(), ri + 1000)


def test_topk_merge(evdr, cuda):
    g = torch.Generator().manual_seed(2)
    cs = torch.randn(5, 80, generator=g)
    cs[:, 10:20] = cs[:, 0:1]
    ci = torch.stack([torch.randperm(10_000, generator=g)[:80] for _ in range(5)])
    ci[0, 3] = -1      # invalid candidate is skipped
    vals, ids = evdr.topk_merge(cs.to(cuda), ci.to(cuda), 10)
    for q in range(5):
        cand = [(-cs[q, j].item(), ci[q, j].item()) for j in range(80) if ci[q, j] >= 0]
        want = [i for _, i in sorted(cand)[:10]]
        assert ids[q].tolist() == want


# ----------------------------------------------------------------------------- metrics
def _random_eval_problem(seed, nq=40, nd=90, graded=True, ties=True):
    rng = np.random.default_rng(seed)
    docs = [f"doc-{rng.integers(0, 10**6):06d}-{i}" for i in range(nd)]
    scores = rng.standard_normal((nq, nd)).astype(np.float32)
    if ties:
        scores = np.round(scores, 1)
    qrels, results = {}, {}
    for q in range(nq):
        qid = f"q{q}"
        rel_docs = rng.choice(nd, size=rng.integers(1, 4), replace=False)
        qrels[qid] = {docs[j]: int(rng.integers(1, 4)) if graded else 1 for j in rel_docs}
        if q % 7 == 0:
            qrels[qid]["never-retrieved"] = 2
        if q % 11 == 0:
            qrels[qid][docs[0]] = 0            # judged non-relevant
        results[qid] = {docs[j]: float(scores[q, j]) for j in range(nd)}
    qrels["unscored-query"] = {docs[0]: 1}
    return qrels, results


@pytest.mark.parametrize("seed,ties", [(0, False), (1, True), (2, True)])
def test_evaluator_matches_trec_oracle(evdr, cuda, seed, ties):
    qrels, results = _random_eval_problem(seed, ties=ties)
    ev = evdr.CustomRetrievalEvaluator()
    got = ev.compute_mteb_metrics(qrels, results)
    want = to.compute_mteb_metrics(qrels, results, ev.k_values)
    assert got.keys() == want.keys()
    for fam in want:
        for key, v in want[fam].items():
            assert got[fam][key] == pytest.approx(v, abs=1.1e-5), (fam, key)
    # the two numbers the reference's scripts consume must be identical after round(., 5)
    assert got["NDCG"]["NDCG@5"] == want["NDCG"]["NDCG@5"]
    assert got["Recall"]["Recall@1"] == want["Recall"]["Recall@1"]


# ----------------------------------------------------------------------------- drop-in module
def test_dropin_module_binds_by_reference_import_path(cuda):
    root = Path(__file__).resolve().parent.parent
    dropin = str(root / "efficient-visual-document-retrieval_b200" / "dropin")
    sys.path.insert(0, dropin)
    try:
        for k in [k for k in sys.modules if k == "evaluator" or k.startswith("evaluator.")]:
            del sys.modules[k]
        mod = importlib.import_module("evaluator.retrieval")
        from evaluator.retrieval import score_multi_vector_masked, CustomRetrievalEvaluator  # noqa: F401
        z = U.load("micro")
        out = score_multi_vector_masked(torch.from_numpy(z["Q"]).to(cuda), torch.from_numpy(z["P"]).to(cuda),
                                        torch.from_numpy(z["qmask"]).to(cuda), torch.from_numpy(z["pmask"]).to(cuda), 64)
        np.testing.assert_allclose(out.cpu().numpy(), z["out"], rtol=U.EXACT_RTOL, atol=U.EXACT_ATOL)
        assert "efficient-visual-document-retrieval_b200" in mod._m.__name__
    finally:
        sys.path.remove(dropin)
        for k in [k for k in sys.modules if k == "evaluator" or k.startswith("evaluator.")]:
            del sys.modules[k]


@pytest.mark.parametrize("nq,n", [(40, 300), (3, 150_000)])
def test_fused_topk_metrics(evdr, cuda, nq, n):
    """evdr_topk_metrics (one pass) == evdr_topk + evdr_retrieval_metrics == the trec oracle."""
    g = torch.Generator().manual_seed(nq + n)
    s = torch.randn(nq, n, generator=g).round(decimals=2)              # ties
    off, ids, rels = [0], [], []
    for q in range(nq):
        for j in torch.randperm(n, generator=g)[: int(torch.randint(1, 5, (1,), generator=g))].tolist():
            ids.append(j)
            rels.append(int(torch.randint(0, 4, (1,), generator=g)))
        off.append(len(ids))
    off_t, ids_t, rels_t = torch.tensor(off), torch.tensor(ids), torch.tensor(rels)
    ks = [1, 3, 5, 10, 50, 70, 100]
    ts, ti, m = evdr.topk_metrics(s.to(cuda), 100, off_t, ids_t, rels_t, ks)
    ts2, ti2 = evdr.top_k(s.to(cuda), 100)
    m2 = evdr.retrieval_metrics(ti2, off_t, ids_t, rels_t, ks)
    assert torch.equal(ti, ti2) and torch.equal(ts, ts2) and torch.equal(m, m2)
    m = m.cpu().numpy()
    for q in range(min(nq, 10)):
        qrel = {f"{j:07d}": r for j, r in zip(ids[off[q]:off[q + 1]], rels[off[q]:off[q + 1]])}
        ranking = [f"{j:07d}" for j in ti[q].tolist()]
        want = to.per_query_measures(qrel, ranking, ks)
        for jk, k in enumerate(ks):
            assert m[0, jk, q] == pytest.approx(want[f"ndcg@{k}"], abs=1e-12)
            assert m[1, jk, q] == pytest.approx(want[f"map@{k}"], abs=1e-12)
            assert m[2, jk, q] == pytest.approx(want[f"recall@{k}"], abs=1e-12)
            assert m[3, jk, q] == pytest.approx(want[f"p@{k}"], abs=1e-12)


def test_evaluate_scores_fast_path(evdr, cuda):
    """Device-resident evaluation (no dicts): same numbers as the dict signature."""
    qrels, results = _random_eval_problem(5, nq=30, nd=120, ties=True)
    qids = [q for q in results]
    docs = list(results[qids[0]].keys())
    dense = torch.tensor([[results[q][d] for d in docs] for q in qids], dtype=torch.float32)
    order = sorted(docs, reverse=True)
    tie = torch.empty(len(docs), dtype=torch.int32)
    tie[torch.tensor([docs.index(d) for d in order])] = torch.arange(len(docs), dtype=torch.int32)
    off, ids, rels, extra = [0], [], [], len(docs)
    for q in qids:
        for d, r in qrels[q].items():
            if d in docs:
                ids.append(docs.index(d))
            else:
                ids.append(extra)
                extra += 1
            rels.append(r)
        off.append(len(ids))
    ev = evdr.CustomRetrievalEvaluator()
    got = ev.evaluate_scores(dense.to(cuda), torch.tensor(off), torch.tensor(ids), torch.tensor(rels), tie_rank=tie.to(cuda),
                             n_judged=len(qrels))
    want = ev.compute_mteb_metrics(qrels, results)
    assert got == want
    # without n_judged the MRR denominator is the number of scored rows; everything else is unchanged
    got_scored = ev.evaluate_scores(dense.to(cuda), torch.tensor(off), torch.tensor(ids), torch.tensor(rels),
                                    tie_rank=tie.to(cuda))
    assert {k: v for k, v in got_scored.items() if k != "mRR"} == {k: v for k, v in want.items() if k != "mRR"}


def test_from_object_arrays_matches_padded_pipeline(evdr, cuda):
    """The on-disk layout (object arrays of (Li,128) + attention / image masks) packed directly equals
    the reference's route: zero-pad to (N, Lmax, 128), mask = valid & attn & img
    (utils/preprocess_data.py:80-105), then pack."""
    rng = np.random.default_rng(0)
    lens = [37, 1, 200, 64, 16, 129, 5]
    docs = np.empty(len(lens), dtype=object)
    attn = np.empty(len(lens), dtype=object)
    img = np.empty(len(lens), dtype=object)
    for i, L in enumerate(lens):
        x = rng.standard_normal((L, 128)).astype(np.float32)
        docs[i] = x / np.linalg.norm(x, axis=-1, keepdims=True)
        attn[i] = rng.random(L) > 0.1
        img[i] = (rng.random(L) > 0.2).astype(np.int64)
    attn[3] = attn[3][:40]            # a mask shorter than its page: the rest is padded False
    img[4] = np.zeros(16, dtype=bool)  # a page with no valid token
    pk = evdr.PackedPages.from_object_arrays(docs, attn, img, device=cuda)
    # reference route restated with numpy (pad_tokens_object / pad_mask_object)
    Lmax = max(lens)
    P = np.zeros((len(lens), Lmax, 128), np.float32)
    valid = np.zeros((len(lens), Lmax), bool)
    am = np.zeros_like(valid)
    im = np.zeros_like(valid)
    for i, L in enumerate(lens):
        P[i, :L] = docs[i]
        valid[i, :L] = True
        a = np.asarray(attn[i]).astype(bool)
        am[i, :min(Lmax, a.shape[0])] = a[:Lmax]
        b = np.asarray(img[i]).astype(bool)
        im[i, :min(Lmax, b.shape[0])] = b[:Lmax]
    pm = valid & am & im
    ref = evdr.PackedPages.from_padded(torch.from_numpy(P).to(cuda), torch.from_numpy(pm).to(cuda), varlen=True)
    assert torch.equal(pk.chunk_off, ref.chunk_off) and pk.total_chunks == ref.total_chunks
    assert torch.equal(pk.tokens, ref.tokens)
    Q = torch.nn.functional.normalize(torch.randn(6, 20, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(cuda)
    qm = torch.ones(6, 20, dtype=torch.bool, device=cuda)
    want = mo.maxsim_masked(Q.cpu().numpy(), P, qm.cpu().numpy(), pm)
    U.assert_scores_close(pk.score(Q, qm).cpu().numpy(), want, rel=2e-3)   # operands not bf16-representable
    pk3 = evdr.PackedPages.from_object_arrays(docs, attn, img, device=cuda, k_blocks=3)
    U.assert_scores_close(pk3.score(Q, qm).cpu().numpy(), want, rel=1e-5)
