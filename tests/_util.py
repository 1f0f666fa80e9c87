"""Shared helpers for the tests: golden fixtures, seeded input regeneration, tolerances."""
from __future__ import annotations

import ast
import hashlib
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"

# north_star: scores within 1e-3 relative with fp32 accumulation.  Scores are sums of <= 32 cosines,
# so the bf16 path is compared with |a-b| <= REL_TOL * max(|ref|, 1).
REL_TOL = 1e-3
# exact fp32 kernels differ from the reference only by summation order
EXACT_RTOL, EXACT_ATOL = 2e-5, 2e-5


def load(name: str):
    return np.load(GOLDEN / f"{name}.npz", allow_pickle=False)


def bf16_from_i16(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.copy()).view(torch.bfloat16)


def sha(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        a = t.detach().cpu()
        if a.dtype == torch.bfloat16:
            a = a.view(torch.int16)
        h.update(np.ascontiguousarray(a.numpy()).tobytes())
    return h.hexdigest()


def seeded_case(name: str):
    """Regenerate the inputs of a seeded golden case on the CPU and check their digest."""
    from evdr_b200 import synth
    z = load("seeded")
    npg, lp, nq, lq, pk, qk = ast.literal_eval(str(z[f"{name}/spec"]))
    P, pm = synth.make_pages(npg, lp, **pk)
    Q, qm, rel = synth.make_queries(P, pm, nq, lq, **qk)
    assert sha(Q, P, qm, pm) == str(z[f"{name}/sha"]), f"synthetic generator drifted for case {name}"
    return Q, P, qm, pm, rel, z


SEEDED_CASES = ["full1030", "comp64", "varlen", "lq1", "student_mf5"]


def assert_scores_close(got, ref, rel=REL_TOL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() <= rel, f"max relative error {err.max():.3e} > {rel:.1e} at {np.unravel_index(err.argmax(), err.shape)}"
