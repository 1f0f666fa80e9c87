"""GPU parity of the score-matrix consumers (losses.cu) against the golden outputs of the unmodified
reference (tests/golden/losses.npz: criterion.py losses + autograd, rank-gap block of
mainv3_iter_liscore_QA_hardtoken.py:368-385) and against the float64 oracle on other shapes."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo
from tests import _util as U

pytestmark = pytest.mark.gpu

CASES = ["b32_n500_k10", "b32_n1000_k40_T05", "b5_n70_k40_T2", "b1_n12_k12"]
# fp32 kernels vs the reference's fp32 torch ops: summation order only
LOSS_RTOL, GRAD_RTOL = 2e-5, 2e-4


def _case(name, cuda):
    z = np.load(U.GOLDEN / "losses.npz")
    B, N, k, aux, _ = (int(v) for v in z[f"{name}/spec"])
    s = torch.from_numpy(z[f"{name}/s"]).to(cuda)
    t = torch.from_numpy(z[f"{name}/t"]).to(cuda)
    return z, B, N, k, aux, float(z[f"{name}/T"]), s, t


def _close_grad(got, ref, what):
    np.testing.assert_allclose(got, ref, rtol=GRAD_RTOL, atol=2e-6 * np.abs(ref).max(), err_msg=what)


@pytest.mark.parametrize("name", CASES)
def test_losses_match_reference(evdr, cuda, name):
    z, B, N, k, aux, T, s, t = _case(name, cuda)
    for key, fn in (("listwise", lambda x: evdr.listwise_distillation_loss(x, t, k=k, temperature=T)),
                    ("mse", lambda x: evdr.score_preserving_loss(x, t)),
                    ("pairwise", lambda x: evdr.pairwise_distillation_loss(x, t)),
                    ("lambda", lambda x: evdr.lambda_loss(x, t, alpha=1.0 if T == 1.0 else 0.5)),
                    ("ranknce", lambda x: evdr.ranknce_loss(x, t, temperature=T, lambda_weight=0.7)),
                    ("infonce_distill", lambda x: evdr.infonce_distillation_loss(x, t, temperature=0.07))):
        sg = s.clone().requires_grad_(True)
        loss = fn(sg)
        assert loss.shape == () and loss.dtype == torch.float32 and loss.device == s.device
        loss.backward()
        if key == "lambda":
            # MaxSim scores are ~10-30 where sigmoid(t) = 1 - O(1e-6): the gain |sig(t_i) - sig(t_j)| is a
            # small multiple of the fp32 spacing, and 1/log2(r+2) of neighbouring ranks nearly cancel, so a
            # last-bit difference between expf / log2f implementations moves single terms by percents; the
            # loss (a sum over B*N*N/2 terms) and the gradient agree to ~1e-4
            np.testing.assert_allclose(loss.item(), z[f"{name}/{key}_loss"], rtol=1e-3, err_msg=key)
            ref_g = z[f"{name}/{key}_grad"]
            np.testing.assert_allclose(sg.grad.cpu().numpy(), ref_g, rtol=1e-2, atol=1e-3 * np.abs(ref_g).max())
            continue
        np.testing.assert_allclose(loss.item(), z[f"{name}/{key}_loss"], rtol=LOSS_RTOL, err_msg=key)
        _close_grad(sg.grad.cpu().numpy(), z[f"{name}/{key}_grad"], key)


@pytest.mark.parametrize("name", CASES)
def test_liscore_combination_is_one_launch_and_linear(evdr, cuda, name):
    """The *_liscore recipes add lambda_list * listwise + lambda_score * mse (mainv2_iter_liscore.py:
    301-303): both come from one evdr_liscore_loss call and the backward combines the two gradients."""
    z, B, N, k, aux, T, s, t = _case(name, cuda)
    sg = s.clone().requires_grad_(True)
    ll, lm = evdr.liscore_losses(sg, t, k=k, temperature=T)
    (0.7 * ll + 0.25 * lm).backward()
    want = 0.7 * z[f"{name}/listwise_grad"] + 0.25 * z[f"{name}/mse_grad"]
    _close_grad(sg.grad.cpu().numpy(), want, "combined")
    # bit-reproducible
    sg2 = s.clone().requires_grad_(True)
    ll2, lm2 = evdr.liscore_losses(sg2, t, k=k, temperature=T)
    (0.7 * ll2 + 0.25 * lm2).backward()
    assert torch.equal(ll, ll2) and torch.equal(lm, lm2) and torch.equal(sg.grad, sg2.grad)


@pytest.mark.parametrize("name", CASES)
def test_rank_gap_matches_reference(evdr, cuda, name):
    z, B, N, k, aux, T, s, t = _case(name, cuda)
    rg = evdr.rank_gap(t, s, k=k, aux_docs=aux)
    for got, key in ((rg.rank_t, "rank_t"), (rg.rank_s, "rank_s"), (rg.gap, "gap"), (rg.G, "G"),
                     (rg.topk_idx, "topk_idx")):
        np.testing.assert_array_equal(got.cpu().numpy(), z[f"{name}/{key}"], err_msg=key)   # bit-exact
    gap = np.abs(z[f"{name}/gap"])
    rows = np.arange(B)[:, None]
    got_aux = rg.aux_doc_idx.cpu().numpy()
    assert got_aux.shape == z[f"{name}/aux_doc_idx_q"].shape
    # same |gap| values in the same order; equal gaps may name another page of the teacher's top-k
    np.testing.assert_array_equal(gap[rows, got_aux], gap[rows, z[f"{name}/aux_doc_idx_q"]])
    assert all(set(got_aux[b]) <= set(z[f"{name}/topk_idx"][b]) for b in range(B))


@pytest.mark.parametrize("B,N,k,T", [(1, 1, 1, 1.0), (3, 257, 100, 0.07), (64, 2048, 10, 1.0), (1663, 277, 10, 1.0),
                                     (300, 277, 40, 0.5)])
def test_losses_vs_oracle_other_shapes(evdr, cuda, B, N, k, T):
    g = torch.Generator().manual_seed(B + N)
    t = 12.0 + 3.0 * torch.randn(B, N, generator=g)
    s = t + 1.5 * torch.randn(B, N, generator=g)
    sd, td = s.to(cuda), t.to(cuda)
    pairs = [("listwise", lambda x: evdr.listwise_distillation_loss(x, td, k=k, temperature=T),
              lo.listwise_distillation_loss(s.numpy(), t.numpy(), k, T)),
             ("mse", lambda x: evdr.score_preserving_loss(x, td), lo.score_preserving_loss(s.numpy(), t.numpy()))]
    pairs.append(("ranknce", lambda x: evdr.ranknce_loss(x, td, temperature=T, lambda_weight=2.0),
                  lo.ranknce_loss(s.numpy(), t.numpy(), T, 2.0)))
    if B * N * N <= 60_000_000:
        pairs.append(("pairwise", lambda x: evdr.pairwise_distillation_loss(x, td),
                      lo.pairwise_distillation_loss(s.numpy(), t.numpy())))
        pairs.append(("lambda", lambda x: evdr.lambda_loss(x, td, alpha=2.0),
                      lo.lambda_loss(s.numpy(), t.numpy(), alpha=2.0)))
    for key, fn, (ref_loss, ref_grad) in pairs:
        sg = sd.clone().requires_grad_(True)
        loss = fn(sg)
        loss.backward()
        lr, gr, ga = (1e-3, 1e-2, 1e-3) if key == "lambda" else (5e-5, 1e-3, 5e-6)   # see test_losses_match_reference
        np.testing.assert_allclose(loss.item(), ref_loss, rtol=lr, atol=1e-7, err_msg=key)
        np.testing.assert_allclose(sg.grad.cpu().numpy(), ref_grad, rtol=gr,
                                   atol=ga * max(np.abs(ref_grad).max(), 1e-30), err_msg=key)


def test_rank_gap_ties_keep_index_order_and_strided_rows(evdr, cuda):
    t = torch.tensor([[1.0, 3.0, 3.0, 0.5, 3.0], [2.0, 2.0, 2.0, 2.0, 2.0]], device=cuda)
    s = torch.tensor([[5.0, 4.0, 3.0, 2.0, 1.0], [1.0, 2.0, 3.0, 4.0, 5.0]], device=cuda)
    wide = torch.zeros(2, 9, device=cuda)
    wide[:, :5] = t
    rg = evdr.rank_gap(wide[:, :5], s, k=3, aux_docs=2)          # non-contiguous rows (ld = 9)
    ref = lo.rank_gap(t.cpu().numpy(), s.cpu().numpy(), 3, 2)
    assert rg.rank_t.tolist() == [[3, 0, 1, 4, 2], [0, 1, 2, 3, 4]]
    for key in ("rank_t", "rank_s", "gap", "G", "topk_idx", "aux_doc_idx"):
        np.testing.assert_array_equal(getattr(rg, key).cpu().numpy(), ref[key], err_msg=key)


def test_loss_argument_errors(evdr, cuda):
    s = torch.zeros(2, 5, device=cuda)
    with pytest.raises(ValueError):                       # torch.topk raises when k > N (criterion.py:132)
        evdr.listwise_distillation_loss(s, s, k=6)
    with pytest.raises(ValueError):
        evdr.pairwise_distillation_loss(s, torch.zeros(2, 6, device=cuda))
    with pytest.raises(evdr._lib.EvdrError):              # no CPU fallback
        evdr.score_preserving_loss(s.cpu(), s.cpu())


def test_training_step_with_fused_losses(evdr, cuda):
    """Scorer -> fused liscore loss -> scorer backward: gradient on the student pages equals the one
    obtained with the same losses written in torch (the reference's criterion.py formulas)."""
    import torch.nn.functional as F
    g = torch.Generator(device=cuda).manual_seed(11)
    npg, ls, nq = 200, 41, 32
    Pt, pmt = evdr.synth.make_pages(npg, 206, seed=5, device=cuda, n_holes=3, empty_frac=0.0, min_len=60)
    Q, qm, _ = evdr.synth.make_queries(Pt, pmt, nq, 32, seed=6, sigma=1.0, min_len=8, drop=0.1)
    Pbar = torch.randn(npg, ls, 128, generator=g, device=cuda)
    pms = torch.rand(npg, ls, generator=g, device=cuda) > 0.1
    with torch.no_grad():
        sc_t = evdr.score_multi_vector_masked(Q, Pt, qm, pmt, 32)
    grads = []
    for fused in (True, False):
        param = Pbar.clone().requires_grad_(True)
        Ps = evdr.l2_normalize_masked(param, pms)
        sc_s = evdr.score_multi_vector_masked(Q, Ps, qm, pms, 32)
        if fused:
            ll, lm = evdr.liscore_losses(sc_s, sc_t, k=10, temperature=1.0)
        else:
            lps, pt = F.log_softmax(sc_s, dim=1), F.softmax(sc_t, dim=1)
            _, idx = torch.topk(pt, 10, dim=1)
            ll = -(pt.gather(1, idx) * lps.gather(1, idx)).sum(1).mean()
            lm = F.mse_loss(sc_s, sc_t)
        (1.0 * ll + 0.5 * lm).backward()
        grads.append(param.grad)
    rel = ((grads[0] - grads[1]).norm() / grads[1].norm()).item()
    assert rel < 1e-4, rel


def test_dropin_criterion_module_exports_the_reference_names(evdr, cuda):
    """`from criterion import ...` of every main*.py resolves against dropin/criterion.py."""
    import importlib.util
    from pathlib import Path
    path = Path(evdr.__file__).resolve().parent / "dropin" / "criterion.py"
    spec = importlib.util.spec_from_file_location("criterion_dropin", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    names = ["infonce_supervised_loss", "infonce_distillation_loss", "score_preserving_loss",
             "pairwise_distillation_loss", "listwise_distillation_loss", "lambda_loss", "ranknce_loss"]
    assert sorted(mod.__all__) == sorted(names)
    s = torch.randn(4, 30, device=cuda, requires_grad=True)
    t = torch.randn(4, 30, device=cuda)
    for nm in names:
        fn = getattr(mod, nm)
        loss = fn(s, torch.randint(0, 30, (4,), device=cuda)) if nm == "infonce_supervised_loss" else fn(s, t)
        assert loss.shape == () and bool(torch.isfinite(loss))
        loss.backward()
    assert bool(torch.isfinite(s.grad).all())


@pytest.mark.parametrize("B,N,k,T", [(32, 500, 10, 1.0), (32, 1000, 40, 0.1), (7, 70, 40, 0.5)])
def test_losses_vs_the_reference_criterion_on_this_gpu(evdr, cuda, B, N, k, T):
    """All seven losses and their gradients against the reference's OWN criterion.py (unmodified file staged in
    baseline/_ref) evaluated with torch autograd on the same GPU, at MaxSim score magnitudes."""
    from baseline import ref as reference
    if not reference.available():
        pytest.skip(reference.why_unavailable())
    crit = reference.criterion_module()
    g = torch.Generator().manual_seed(B * N)
    t = (14.0 + 2.5 * torch.randn(B, N, generator=g)).to(cuda)
    s0 = (t.cpu() + 1.0 * torch.randn(B, N, generator=g)).to(cuda)
    labels = torch.randint(0, N, (B,), generator=g).to(cuda)
    cases = [("listwise", lambda m, x: m.listwise_distillation_loss(x, t, k=k, temperature=T)),
             ("score_preserving", lambda m, x: m.score_preserving_loss(x, t)),
             ("pairwise", lambda m, x: m.pairwise_distillation_loss(x, t)),
             ("ranknce", lambda m, x: m.ranknce_loss(x, t, temperature=T, lambda_weight=0.7)),
             ("infonce_distillation", lambda m, x: m.infonce_distillation_loss(x, t, temperature=0.07)),
             ("infonce_supervised", lambda m, x: m.infonce_supervised_loss(x, labels, temperature=0.07)),
             ("lambda", lambda m, x: m.lambda_loss(x, t))]
    for name, fn in cases:
        xs, xr = s0.clone().requires_grad_(True), s0.clone().requires_grad_(True)
        lo_, lr_ = fn(evdr, xs), fn(crit, xr)
        lo_.backward()
        lr_.backward()
        gr = xr.grad.cpu().numpy()
        if name == "lambda":      # quantised sigmoid differences (see test_losses_match_reference)
            np.testing.assert_allclose(lo_.item(), lr_.item(), rtol=1e-3, err_msg=name)
            np.testing.assert_allclose(xs.grad.cpu().numpy(), gr, rtol=1e-2, atol=1e-3 * np.abs(gr).max(), err_msg=name)
        else:
            np.testing.assert_allclose(lo_.item(), lr_.item(), rtol=5e-5, atol=1e-7, err_msg=name)
            np.testing.assert_allclose(xs.grad.cpu().numpy(), gr, rtol=1e-3, atol=5e-6 * max(np.abs(gr).max(), 1e-30), err_msg=name)
