"""Seeded random shapes through the drop-in call against the exact fp32 kernel (and, for a corner of every
case, the numpy oracle): query counts across every kernel flavour boundary (4 | 5, 8 | 9, 60 | 61 queries),
page lengths from 1 token to beyond two tiles, arbitrary masks, empty pages and queries, both pack plans."""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as mo
from tests import _util as U

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("plain_bf16")]

_rng = np.random.default_rng(20261018)
_NQ = [1, 2, 4, 5, 7, 8, 9, 12, 15, 16, 17, 31, 32, 33, 47, 60, 61, 64, 65, 99, 128, 140]
_LP = [1, 2, 15, 16, 17, 31, 33, 48, 63, 64, 65, 100, 191, 192, 193, 255, 256, 257, 400, 511, 513, 1030, 1100]
CASES = [(int(_rng.choice(_NQ)), int(_rng.integers(1, 260)), int(_rng.choice(_LP)), int(_rng.integers(1, 33)),
          bool(_rng.integers(0, 2)), int(_rng.integers(0, 1 << 30))) for _ in range(48)]


@pytest.mark.parametrize("nq,npg,lp,lq,varlen,seed", CASES)
def test_random_shapes_vs_exact_kernel(evdr, cuda, nq, npg, lp, lq, varlen, seed):
    g = torch.Generator(device=cuda).manual_seed(seed)
    P = torch.nn.functional.normalize(torch.randn(npg, lp, 128, generator=g, device=cuda), dim=-1).bfloat16()
    Q = torch.nn.functional.normalize(torch.randn(nq, lq, 128, generator=g, device=cuda), dim=-1).bfloat16()
    pm = torch.rand(npg, lp, generator=g, device=cuda) > 0.25
    qm = torch.rand(nq, lq, generator=g, device=cuda) > 0.2
    pm[torch.rand(npg, generator=g, device=cuda) < 0.05] = False          # empty pages
    qm[torch.rand(nq, generator=g, device=cuda) < 0.05] = False            # empty queries
    if npg > 3:
        pm[2] = True                                                        # a page with every slot valid
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    got = evdr.score_multi_vector_masked(Q, P, qm.long(), pm.float(), chunk_p=7)      # masks of any dtype
    U.assert_scores_close(got.cpu().numpy(), ref.cpu().numpy())
    packed = evdr.PackedPages.from_padded(P, pm, varlen=varlen).score(Q, qm)
    U.assert_scores_close(packed.cpu().numpy(), ref.cpu().numpy())
    empty_p, empty_q = ~pm.any(1), ~qm.any(1)
    assert (got[:, empty_p] == 0).all() and (got[empty_q] == 0).all()
    a, b = min(nq, 3), min(npg, 4)
    o = mo.maxsim_masked(Q[:a].float().cpu().numpy(), P[:b].float().cpu().numpy(), qm[:a].cpu().numpy(),
                         pm[:b].cpu().numpy())
    U.assert_scores_close(got[:a, :b].cpu().numpy(), o)


@pytest.mark.parametrize("nq,npg,ls,seed", [(int(_rng.integers(1, 40)), int(_rng.integers(1, 150)),
                                             int(_rng.choice([1, 5, 16, 21, 41, 103, 206])), int(_rng.integers(0, 1 << 30)))
                                            for _ in range(12)])
def test_random_gradient_calls_vs_exact_kernel(evdr, cuda, nq, npg, ls, seed):
    """Tensor-core search + fp32 rescoring vs the exact fp32 kernels on student-sized pages: same scores, the
    same gradients apart from exact-tie winners, zero gradient on masked tokens."""
    g = torch.Generator(device=cuda).manual_seed(seed)
    P = torch.randn(npg, ls, 128, generator=g, device=cuda)
    pm = torch.rand(npg, ls, generator=g, device=cuda) > 0.2
    P = evdr.l2_normalize_masked(P, pm)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g, device=cuda), dim=-1)
    qm = torch.rand(nq, 32, generator=g, device=cuda) > 0.15
    gout = torch.randn(nq, npg, generator=g, device=cuda)
    outs, grads = [], []
    for prec in (None, "fp32"):
        Pg = P.clone().requires_grad_(True)
        o = evdr.score_multi_vector_masked(Q, Pg, qm, pm, 32, precision=prec)
        o.backward(gout)
        outs.append(o.detach())
        grads.append(Pg.grad)
    np.testing.assert_allclose(outs[0].cpu().numpy(), outs[1].cpu().numpy(), rtol=2e-5, atol=2e-5)
    assert grads[0][~pm].abs().max() == 0 if (~pm).any() else True
    denom = grads[1].norm().item()
    if denom > 0:
        assert ((grads[0] - grads[1]).norm() / denom).item() < 2e-2
