"""Hard-token mining on device.

Reference: the Python double loop of mainv3_iter_liscore_QA_hardtoken.py:398-425 -- for every query of
the batch and each of its `aux_docs` hard pages it builds ``q_tokens @ doc_tok.T``, masks invalid page
tokens with -inf, takes ``sim.max(dim=0)`` then ``argmax`` and reads the index back with ``.tolist()``
(one host sync per query).  Here: one kernel launch for the whole batch, no sync.
"""
from __future__ import annotations

import torch

from . import _lib
from .corpus import _float_operand, _mask_u8


def hard_token_argmax(Q: torch.Tensor, qmask: torch.Tensor, P: torch.Tensor, pmask: torch.Tensor,
                      doc_idx: torch.Tensor) -> torch.Tensor:
    """Q (B, Lq, D), P (N, Lp, D), doc_idx (B, A) int64 -> best page-token index (B, A) int64:
    argmax_m max_{valid n} <Q[b,n], P[doc_idx[b,a], m]> over valid page tokens, first index on ties."""
    _lib.require_cuda(Q, "Q")
    _lib.require_cuda(P, "P")
    Qc, Pc = _float_operand(Q), _float_operand(P)
    qm, pm = _mask_u8(qmask), _mask_u8(pmask)
    B, lq, d = Qc.shape
    N, lp, _ = Pc.shape
    di = doc_idx.to(device=Q.device, dtype=torch.int64).contiguous()
    if di.dim() != 2 or di.shape[0] != B:
        raise ValueError("doc_idx must be (B, A)")
    A = di.shape[1]
    out = torch.empty((B, A), dtype=torch.int64, device=Q.device)
    if A == 0:
        return out
    lib = _lib.load()
    _lib.check(lib.evdr_hard_token_argmax(_lib.handle(Q.device), Qc.data_ptr(), _lib.dtype_code(Qc), Pc.data_ptr(),
                                          _lib.dtype_code(Pc), qm.data_ptr(), pm.data_ptr(), di.data_ptr(), B, A, lq,
                                          lp, d, out.data_ptr(), _lib.stream_ptr(Q.device)),
               "evdr_hard_token_argmax")
    return out


def hard_tokens(Q, qmask, P, pmask, doc_idx) -> torch.Tensor:
    """The mined tokens themselves, (B*A, 1, D): the 1-token virtual queries of :427-429 (before the
    optional noise / re-normalisation, which stay with the caller)."""
    idx = hard_token_argmax(Q, qmask, P, pmask, doc_idx)
    tok = P[doc_idx.to(P.device).reshape(-1), idx.reshape(-1)]
    return tok.unsqueeze(1)
