"""Golden vectors for the score-matrix consumers, by EXECUTING the unmodified reference (build
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_golden_losses.py        # writes tests/golden/losses.npz

* ``criterion.py`` is imported from /root/reference as is; losses and their autograd gradients w.r.t.
  the student scores are stored.
* The rank-gap block lives inside a training-step function (mainv3_iter_liscore_QA_hardtoken.py:
  368-385), so those source lines are read from the reference file at run time, dedented and
  ``exec``-ed on the stored inputs -- nothing of the reference is copied into this repository.

Inputs: seeded score matrices shaped like a training step (32 x 500 / 32 x 1000, student = teacher +
noise so that ranks differ but correlate), a small 5 x 70 case (tabfquad has 70 pages: k = 40 < N),
and temperatures 1.0 / 0.5 / 2.0.
"""
from __future__ import annotations

import importlib
import sys
import textwrap
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")

CASES = {  # name: (B, N, k, temperature, aux_docs, seed)
    "b32_n500_k10": (32, 500, 10, 1.0, 4, 1),
    "b32_n1000_k40_T05": (32, 1000, 40, 0.5, 4, 2),
    "b5_n70_k40_T2": (5, 70, 40, 2.0, 3, 3),
    "b1_n12_k12": (1, 12, 12, 1.0, 12, 4),
}


def make_scores(B, N, seed):
    """Tie-free rows: torch.argsort leaves the order of equal scores unspecified, so a fixture with a
    tie would pin an accident of the sort implementation.  The seed is advanced by 1000 until no row of
    either matrix holds a repeated value (fp32 collisions do occur at N = 500)."""
    while True:
        g = torch.Generator().manual_seed(seed)
        t = 12.0 + 3.0 * torch.randn(B, N, generator=g)
        s = t + 1.5 * torch.randn(B, N, generator=g)
        if all(len(torch.unique(r)) == N for m in (s, t) for r in m):
            return s, t
        seed += 1000


def main():
    sys.path.insert(0, str(REF))
    crit = importlib.import_module("criterion")
    assert Path(crit.__file__).resolve().is_relative_to(REF), crit.__file__
    src = (REF / "mainv3_iter_liscore_QA_hardtoken.py").read_text().splitlines()
    block = textwrap.dedent("\n".join(src[368:385]))        # lines 369-385: body of `with torch.no_grad():`
    assert block.lstrip().startswith("rank_t = torch.argsort"), block[:80]
    out = {}
    for name, (B, N, k, T, aux, seed) in CASES.items():
        s, t = make_scores(B, N, seed)
        out[f"{name}/spec"] = np.array([B, N, k, aux, seed], dtype=np.int64)
        out[f"{name}/T"] = np.float64(T)
        out[f"{name}/s"] = s.numpy()
        out[f"{name}/t"] = t.numpy()
        for fn, key, kw in ((crit.listwise_distillation_loss, "listwise", {"k": k, "temperature": T}),
                            (crit.score_preserving_loss, "mse", {}),
                            (crit.pairwise_distillation_loss, "pairwise", {}),
                            (crit.lambda_loss, "lambda", {"alpha": 1.0 if T == 1.0 else 0.5}),
                            (crit.ranknce_loss, "ranknce", {"temperature": T, "lambda_weight": 0.7}),
                            (crit.infonce_distillation_loss, "infonce_distill", {"temperature": 0.07})):
            sg = s.clone().requires_grad_(True)
            loss = fn(sg, t, **kw)
            loss.backward()
            out[f"{name}/{key}_loss"] = loss.detach().numpy()
            out[f"{name}/{key}_grad"] = sg.grad.numpy()
        env = {"torch": torch, "sc_t": t, "sc_s": s, "GAP_TOPK_RETURN": 20, "k": k, "aux_docs": aux,
               "Qb": torch.zeros(B, 1), "device": torch.device("cpu")}
        exec(block, env)
        for key in ("rank_t", "rank_s", "gap", "G", "gap_top_val", "gap_top_idx", "topk_idx", "aux_doc_idx_q"):
            out[f"{name}/{key}"] = env[key].numpy()
    np.savez_compressed(HERE / "losses.npz", **out)
    print("wrote", HERE / "losses.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
