"""Pure-Python restatement of the retrieval metrics the reference obtains from
mteb -> pytrec_eval (TEST INFRASTRUCTURE; PARITY UNPINNED, see oracle/__init__.py).

Reference call sites: evaluator/retrieval.py:230-255 (CustomRetrievalEvaluator.
compute_mteb_metrics) with k_values [1,3,5,10,50,70,100] (:223); callers read only
metrics["NDCG"]["NDCG@5"] and metrics["Recall"]["Recall@1"]
(mainv2_iter_listwise.py:375-394).

Third-party code restated here (absent from /root/reference, unpinned there):
* trec_eval 9.0 measures as exposed by pytrec_eval: ``ndcg_cut_k``, ``map_cut_k``,
  ``recall_k``, ``P_k``.  Ranking = score descending, ties broken by document id
  string DESCENDING (trec_eval sorts results by (sim desc, docno desc)).
  ndcg uses LINEAR gain (the rel value) and 1/log2(rank+1) discounts; the ideal
  ranking is the query's judged rels sorted descending.  map_cut divides by the
  number of relevant (rel>0) judged documents; recall likewise; P_k divides by k.
* mteb 1.x ``RetrievalEvaluator.evaluate``: only queries present in both qrels and
  results are scored; each measure is averaged over scored queries and rounded to
  5 decimals.  ``evaluate_custom(..., "mrr")``: per query sort (stable, Python
  ``sorted(..., reverse=True)``) by score, take top k_max, reciprocal rank of the
  first doc with rel>0 within top-k; sum / len(qrels), round 5.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

K_VALUES = [1, 3, 5, 10, 50, 70, 100]  # evaluator/retrieval.py:223


def trec_rank(doc_scores: Dict[str, float]) -> List[str]:
    """Document ids ordered as trec_eval ranks them: score desc, then docid desc."""
    return [d for d, _ in sorted(doc_scores.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)]


def dcg_at_k(gains: Sequence[float], k: int) -> float:
    return sum(g / math.log2(i + 2) for i, g in enumerate(gains[:k]))


def per_query_measures(qrel: Dict[str, int], ranking: Sequence[str], k_values=K_VALUES):
    """All cut-off measures for one query given its ranked doc ids."""
    rels = [qrel.get(d, 0) for d in ranking]
    ideal = sorted((r for r in qrel.values() if r > 0), reverse=True)
    n_rel = len(ideal)
    out = {}
    for k in k_values:
        gains = [float(r) if r > 0 else 0.0 for r in rels[:k]]
        idcg = dcg_at_k([float(r) for r in ideal], k)
        out[f"ndcg@{k}"] = dcg_at_k(gains, k) / idcg if idcg > 0 else 0.0
        hits = 0
        ap = 0.0
        for i, r in enumerate(rels[:k]):
            if r > 0:
                hits += 1
                ap += hits / (i + 1)
        out[f"map@{k}"] = ap / n_rel if n_rel > 0 else 0.0
        out[f"recall@{k}"] = hits / n_rel if n_rel > 0 else 0.0
        out[f"p@{k}"] = hits / k
    return out


def evaluate(relevant_docs: Dict[str, Dict[str, int]],
             results: Dict[str, Dict[str, float]],
             k_values=K_VALUES):
    """Returns (ndcg, map, recall, precision) dicts keyed like mteb:
    'NDCG@k', 'MAP@k', 'Recall@k', 'P@k' -- mean over scored queries, round(…,5)."""
    qids = [q for q in results if q in relevant_docs]
    acc = {}
    for q in qids:
        m = per_query_measures(relevant_docs[q], trec_rank(results[q]), k_values)
        for key, v in m.items():
            acc[key] = acc.get(key, 0.0) + v
    n = max(len(qids), 1)
    ndcg = {f"NDCG@{k}": round(acc.get(f"ndcg@{k}", 0.0) / n, 5) for k in k_values}
    _map = {f"MAP@{k}": round(acc.get(f"map@{k}", 0.0) / n, 5) for k in k_values}
    recall = {f"Recall@{k}": round(acc.get(f"recall@{k}", 0.0) / n, 5) for k in k_values}
    prec = {f"P@{k}": round(acc.get(f"p@{k}", 0.0) / n, 5) for k in k_values}
    return ndcg, _map, recall, prec


def mrr(relevant_docs, results, k_values=K_VALUES):
    """mteb 1.x evaluate_custom(..., 'mrr'): stable sort by score desc (insertion order on ties)."""
    k_max = max(k_values)
    tot = {k: 0.0 for k in k_values}
    for q, doc_scores in results.items():
        top = sorted(doc_scores.items(), key=lambda kv: kv[1], reverse=True)[:k_max]
        rel = {d for d, r in relevant_docs.get(q, {}).items() if r > 0}
        for k in k_values:
            for rank, (d, _) in enumerate(top[:k]):
                if d in rel:
                    tot[k] += 1.0 / (rank + 1)
                    break
    # BEIR / mteb 1.x divide by len(qrels) -- every judged query counts, scored or not (the other measures
    # average over the scored queries only); identical whenever every judged query has results, as in the
    # reference's scripts
    n = max(len(relevant_docs), 1)
    return {f"MRR@{k}": round(tot[k] / n, 5) for k in k_values}


def compute_mteb_metrics(relevant_docs, results, k_values=K_VALUES):
    """Same return structure as CustomRetrievalEvaluator.compute_mteb_metrics (:230-255)."""
    ndcg, _map, recall, prec = evaluate(relevant_docs, results, k_values)
    return {"NDCG": ndcg, "mAP": _map, "Recall": recall, "Precision": prec,
            "mRR": mrr(relevant_docs, results, k_values)}
