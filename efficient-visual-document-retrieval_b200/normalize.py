"""Fused mask * l2_normalize with autograd.

Reference: ``l2_normalize(Pbar_param * pmask_student.unsqueeze(-1))`` every training step
(mainv2_iter_listwise.py:280) with ``l2_normalize(x) = x / (x.norm(dim=-1, keepdim=True) + 1e-12)``
(utils/preprocess_data.py:8-9): three elementwise / reduction passes and their autograd mirror.  One
kernel each way here (evdr_l2norm_fwd / evdr_l2norm_bwd)."""
from __future__ import annotations

import torch

from . import _lib
from .corpus import _mask_u8


class _L2NormMasked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, mask_u8):
        Xc = X.contiguous()
        d = Xc.shape[-1]
        n_tok = Xc.numel() // d
        Y = torch.empty_like(Xc)
        lib = _lib.load()
        _lib.check(lib.evdr_l2norm_fwd(_lib.handle(X.device), Xc.data_ptr(), _lib.ptr(mask_u8), n_tok, d,
                                       Y.data_ptr(), _lib.stream_ptr(X.device)), "evdr_l2norm_fwd")
        ctx.save_for_backward(Xc, mask_u8)
        return Y

    @staticmethod
    def backward(ctx, gY):
        Xc, mask_u8 = ctx.saved_tensors
        gY = gY.to(torch.float32).contiguous()
        d = Xc.shape[-1]
        n_tok = Xc.numel() // d
        gX = torch.empty_like(Xc)
        lib = _lib.load()
        _lib.check(lib.evdr_l2norm_bwd(_lib.handle(Xc.device), Xc.data_ptr(), _lib.ptr(mask_u8), gY.data_ptr(),
                                       n_tok, d, gX.data_ptr(), _lib.stream_ptr(Xc.device)), "evdr_l2norm_bwd")
        return gX, None


def l2_normalize_masked(X: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
    """y = (x*mask) / (||x*mask||_2 + 1e-12) over the last axis; fp32, CUDA only, differentiable."""
    _lib.require_cuda(X, "X")
    if X.dtype != torch.float32:
        raise TypeError("l2_normalize_masked expects float32")
    if X.shape[-1] % 4 != 0:
        raise NotImplementedError("last dim must be a multiple of 4")
    m = None if mask is None else _mask_u8(mask)
    if m is not None and m.numel() * X.shape[-1] != X.numel():
        raise ValueError("mask must have one entry per token")
    return _L2NormMasked.apply(X, m)


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """Signature of utils/preprocess_data.py:8-9 (eps is fixed at 1e-12 in the kernel)."""
    if eps != 1e-12:
        raise NotImplementedError("the fused kernel uses eps = 1e-12 like the reference's default")
    return l2_normalize_masked(x, None)
