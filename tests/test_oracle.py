"""CPU: pin the oracle against golden vectors produced by executing the reference
(tests/golden/make_golden.py), and cross-check the independent restatements."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as mo
from oracle import trec_oracle as to
from tests import _util as U


def test_micro_cases_match_reference():
    z = U.load("micro")
    out = mo.maxsim_masked(z["Q"], z["P"], z["qmask"], z["pmask"])
    np.testing.assert_allclose(out, z["out"], rtol=1e-6, atol=1e-6)
    assert (out[:, 2] == 0).all(), "page with every token masked scores exactly 0"
    assert (out[1] == 0).all(), "query with every token masked scores exactly 0"
    # bf16 inputs are up-cast (evaluator/retrieval.py:176-177); integer masks accepted (:179-180)
    qb = torch.from_numpy(z["Q"]).bfloat16().float().numpy()
    pb = torch.from_numpy(z["P"]).bfloat16().float().numpy()
    np.testing.assert_allclose(mo.maxsim_masked(qb, pb, z["qmask"], z["pmask"]), z["out_bf16_in"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(mo.maxsim_masked(z["Q"], z["P"], z["qmask"].astype(np.int64), z["pmask"].astype(np.int32)),
                               z["out_int_masks"], rtol=1e-6, atol=1e-6)


def test_micro_backward_matches_autograd():
    z = U.load("micro")
    dP, dQ = mo.maxsim_backward(z["Q"], z["P"], z["qmask"], z["pmask"], z["gout"])
    np.testing.assert_allclose(dP, z["dP"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dQ, z["dQ"], rtol=1e-5, atol=1e-6)
    # ties: gradient goes only to the first (lowest-index) of the two identical tokens
    assert np.abs(z["dP"][~z["pmask"]]).max() == 0, "masked page tokens receive exactly zero gradient"


def test_medium_case_scores_and_grads():
    z = U.load("medium")
    Q = U.bf16_from_i16(z["Q_bf16"]).float().numpy()
    P = U.bf16_from_i16(z["P_bf16"]).float().numpy()
    out = mo.maxsim_masked(Q, P, z["qmask"], z["pmask"])
    np.testing.assert_allclose(out, z["out"], rtol=2e-6, atol=2e-6)
    dP, dQ = mo.maxsim_backward(Q, P, z["qmask"], z["pmask"], z["gout"])
    np.testing.assert_allclose(dP, z["dP"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dQ, z["dQ"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", U.SEEDED_CASES)
def test_seeded_cases(name):
    Q, P, qm, pm, rel, z = U.seeded_case(name)
    out = mo.maxsim_masked(Q.float().numpy(), P.float().numpy(), qm.numpy(), pm.numpy())
    np.testing.assert_allclose(out, z[f"{name}/out"], rtol=2e-6, atol=2e-6)
    t = mo.maxsim_masked_torch_cpu(Q, P, qm, pm, chunk_p=16).numpy()
    np.testing.assert_allclose(t, z[f"{name}/out"], rtol=2e-6, atol=2e-6)


def test_student_backward_seeded():
    Q, P, qm, pm, rel, z = U.seeded_case("student_mf5")
    dP, _ = mo.maxsim_backward(Q.float().numpy(), P.float().numpy(), qm.numpy(), pm.numpy(), z["student_mf5/gout"])
    np.testing.assert_allclose(dP, z["student_mf5/dP"], rtol=1e-5, atol=1e-6)


def test_list_scorer_left_padding_semantics():
    z = U.load("list_scorer")
    qs = np.split(z["qs"], np.cumsum(z["q_lens"])[:-1])
    ps = np.split(z["ps"], np.cumsum(z["p_lens"])[:-1])
    np.testing.assert_allclose(mo.maxsim_list(qs, ps, batch_size=3), z["out_b3"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(mo.maxsim_list(qs, ps, batch_size=128), z["out_b128"], rtol=2e-6, atol=2e-6)
    with pytest.raises(ValueError, match="No queries provided"):
        mo.maxsim_list([], ps)
    with pytest.raises(ValueError, match="No passages provided"):
        mo.maxsim_list(qs, [])


def test_c_oracle_agrees():
    from oracle.build_oracle import build
    lib = C.CDLL(str(build()))
    z = U.load("medium")
    Q = np.ascontiguousarray(U.bf16_from_i16(z["Q_bf16"]).float().numpy())
    P = np.ascontiguousarray(U.bf16_from_i16(z["P_bf16"]).float().numpy())
    qm = np.ascontiguousarray(z["qmask"].astype(np.uint8))
    pm = np.ascontiguousarray(z["pmask"].astype(np.uint8))
    nq, lq, d = Q.shape
    npg, lp, _ = P.shape
    out = np.zeros((nq, npg), dtype=np.float32)
    am = np.zeros((nq, npg, lq), dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.oracle_maxsim_masked(vp(Q), vp(P), vp(qm), vp(pm), nq, npg, lq, lp, d, vp(out), vp(am))
    assert rc == 0
    np.testing.assert_allclose(out, z["out"], rtol=2e-5, atol=2e-5)
    _, am_np = mo.rowmax_argmax(Q, P, pm)
    # argmax agrees wherever the max is not a near-tie
    assert (am == am_np).mean() > 0.999


def test_l2_normalize():
    x = np.random.default_rng(0).standard_normal((5, 7, 16)).astype(np.float32)
    x[2, 3] = 0
    y = mo.l2_normalize(x)
    ref = torch.from_numpy(x)
    ref = (ref / (ref.norm(dim=-1, keepdim=True) + 1e-12)).numpy()
    np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-7)
    assert (y[2, 3] == 0).all()


# ----------------------------------------------------------------------------- metrics (unpinned)
def test_trec_known_answers():
    qrels = {"q1": {"a": 1}, "q2": {"b": 2, "c": 1}, "q3": {"z": 1}}
    results = {
        "q1": {"a": 0.9, "b": 0.5, "c": 0.1},           # hit at rank 1
        "q2": {"a": 0.9, "b": 0.5, "c": 0.7},           # c (rel 1) rank 2, b (rel 2) rank 3
        "q3": {"a": 0.3, "b": 0.2, "c": 0.1},           # relevant doc never retrieved
    }
    m = to.compute_mteb_metrics(qrels, results, k_values=[1, 3, 5])
    dcg2 = 1 / math.log2(3) + 2 / math.log2(4)
    idcg2 = 2 / math.log2(2) + 1 / math.log2(3)
    assert m["NDCG"]["NDCG@3"] == round((1.0 + dcg2 / idcg2 + 0.0) / 3, 5)
    assert m["NDCG"]["NDCG@1"] == round((1.0 + 0.0 + 0.0) / 3, 5)
    assert m["Recall"]["Recall@1"] == round((1.0 + 0.0 + 0.0) / 3, 5)
    assert m["Recall"]["Recall@3"] == round((1.0 + 1.0 + 0.0) / 3, 5)
    assert m["Precision"]["P@3"] == round((1 / 3 + 2 / 3 + 0.0) / 3, 5)
    assert m["mAP"]["MAP@3"] == round((1.0 + (1 / 2 + 2 / 3) / 2 + 0.0) / 3, 5)
    assert m["mRR"]["MRR@3"] == round((1.0 + 0.5 + 0.0) / 3, 5)


def test_trec_tie_break_is_docid_descending():
    # equal scores: trec_eval ranks the lexicographically LARGER docid first
    assert to.trec_rank({"d1": 0.5, "d3": 0.5, "d2": 0.5, "d0": 0.9}) == ["d0", "d3", "d2", "d1"]


def test_ndcg_cross_check_with_sklearn():
    from sklearn.metrics import ndcg_score
    rng = np.random.default_rng(3)
    for _ in range(20):
        n = 30
        scores = rng.standard_normal(n)
        rels = (rng.random(n) < 0.15).astype(int) * rng.integers(1, 4, n)
        if rels.sum() == 0:
            rels[0] = 1
        docs = [f"d{i:03d}" for i in range(n)]
        qrel = {d: int(r) for d, r in zip(docs, rels) if r > 0}
        ranking = to.trec_rank({d: float(s) for d, s in zip(docs, scores)})
        for k in (1, 5, 10):
            mine = to.per_query_measures(qrel, ranking, [k])[f"ndcg@{k}"]
            sk = ndcg_score(rels[None, :], scores[None, :], k=k)
            assert abs(mine - sk) < 1e-12


def test_map_recall_mrr_cross_check_with_sklearn():
    """Independent implementations (scikit-learn) on tie-free, fully ranked, binary-relevance queries: average
    precision = trec map over the whole ranking, LRAP with a single relevant page = reciprocal rank, recall@k."""
    from sklearn.metrics import average_precision_score, label_ranking_average_precision_score
    rng = np.random.default_rng(11)
    for _ in range(20):
        n = 40
        scores = rng.standard_normal(n)
        rel = np.zeros(n, dtype=int)
        rel[rng.choice(n, int(rng.integers(1, 5)), replace=False)] = 1
        docs = [f"d{i:03d}" for i in range(n)]
        qrel = {d: 1 for d, r in zip(docs, rel) if r}
        res = {d: float(s) for d, s in zip(docs, scores)}
        m = to.per_query_measures(qrel, to.trec_rank(res), [5, n])
        assert abs(m[f"map@{n}"] - average_precision_score(rel, scores)) < 1e-12
        top5 = np.argsort(-scores)[:5]
        assert abs(m["recall@5"] - rel[top5].sum() / rel.sum()) < 1e-12
        single = np.zeros(n, dtype=int)
        single[int(rng.integers(0, n))] = 1
        qrel1 = {docs[int(single.argmax())]: 1}
        rr = to.mrr({"q": qrel1}, {"q": res}, [n])[f"MRR@{n}"]
        assert abs(rr - round(label_ranking_average_precision_score(single[None, :], scores[None, :]), 5)) < 1e-12


# ----------------------------------------------------------------------------------------------
# score-matrix consumers (criterion.py losses, rank-gap mining): oracle vs outputs of the unmodified
# reference stored in tests/golden/losses.npz (tests/golden/make_golden_losses.py)
# ----------------------------------------------------------------------------------------------
LOSS_CASES = ["b32_n500_k10", "b32_n1000_k40_T05", "b5_n70_k40_T2", "b1_n12_k12"]


@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_oracle_matches_reference(name):
    from oracle import loss_oracle as lo
    z = np.load(U.GOLDEN / "losses.npz")
    B, N, k, aux, _ = (int(v) for v in z[f"{name}/spec"])
    T = float(z[f"{name}/T"])
    s, t = z[f"{name}/s"], z[f"{name}/t"]
    for key, (loss, grad) in {"listwise": lo.listwise_distillation_loss(s, t, k, T),
                              "mse": lo.score_preserving_loss(s, t),
                              "pairwise": lo.pairwise_distillation_loss(s, t),
                              "lambda": lo.lambda_loss(s, t, alpha=1.0 if T == 1.0 else 0.5),
                              "ranknce": lo.ranknce_loss(s, t, temperature=T, lambda_weight=0.7)}.items():
        # lambda_loss: the oracle keeps the gain / discount factors in float32 like the reference (their
        # differences are small multiples of the fp32 spacing); what remains is summation order
        lr, gr, ga = (2e-5, 1e-3, 2e-5) if key == "lambda" else (2e-5, 2e-4, 2e-6)
        np.testing.assert_allclose(loss, z[f"{name}/{key}_loss"], rtol=lr, err_msg=key)
        ref_g = z[f"{name}/{key}_grad"]
        np.testing.assert_allclose(grad, ref_g, rtol=gr, atol=ga * np.abs(ref_g).max(), err_msg=key)
    rg = lo.rank_gap(t, s, k, aux)
    for key, ref_key in (("rank_t", "rank_t"), ("rank_s", "rank_s"), ("gap", "gap"), ("G", "G"),
                         ("topk_idx", "topk_idx")):
        np.testing.assert_array_equal(rg[key], z[f"{name}/{ref_key}"], err_msg=key)
    # the hardest `aux` pages: same |gap| values in the same order (equal gaps may name another page)
    gap = np.abs(z[f"{name}/gap"])
    rows = np.arange(B)[:, None]
    np.testing.assert_array_equal(gap[rows, rg["aux_doc_idx"]], gap[rows, z[f"{name}/aux_doc_idx_q"]])


def test_listwise_oracle_rejects_k_above_n():
    from oracle import loss_oracle as lo
    with pytest.raises(ValueError):
        lo.listwise_distillation_loss(np.zeros((2, 5)), np.zeros((2, 5)), k=6)
