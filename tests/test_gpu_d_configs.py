"""GPU parity at the shapes of BASELINE.json's configs against the reference's OWN
``score_multi_vector_masked`` (unmodified evaluator/retrieval.py:166-213 staged under baseline/_ref by
tools/stage_reference.py) run torch-eager on the same device -- the reference's eager path is the bar on a
B200, SURVEY 2.1 -- and a training step through the backward (config 5).  Only if the staged reference is
absent does the test-local restatement below stand in (a warning says so)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import _util as U

from baseline import ref as reference

pytestmark = pytest.mark.gpu


def torch_maxsim(Q, P, qmask, pmask, chunk=32):
    """The comparator: the reference's own function when staged, else the restatement."""
    if reference.available():
        return reference.scorer_module().score_multi_vector_masked(Q, P, qmask, pmask, chunk)
    import warnings
    warnings.warn("baseline/_ref is not staged: comparing with the test-local restatement of the reference")
    return _torch_maxsim_restated(Q, P, qmask, pmask, chunk)


def test_comparator_is_the_unmodified_reference():
    """The staged files are byte-identical to the manifest written when they were copied from the reference."""
    if not reference.available():
        pytest.skip(reference.why_unavailable())
    reference.verify()
    fn = reference.scorer_module().score_multi_vector_masked
    assert "baseline/_ref/evaluator/retrieval.py" in fn.__code__.co_filename.replace("\\", "/")


@pytest.mark.parametrize("dtype,batch_size", [(torch.float32, 128), (torch.float32, 7), (torch.bfloat16, 16)])
def test_A2_list_scorer_vs_the_reference_on_this_gpu(evdr, cuda, dtype, batch_size, capsys):
    """A2 against the reference's OWN BaseVisualRetrieverProcessor.score_multi_vector (evaluator/retrieval.py:101-150,
    left_padding :30-45 hard-codes 'cuda', so it can only run on the GPU box): lists of ragged (Li, 128) tensors,
    blocks of batch_size zero-LEFT-padded pages whose zero tokens take part in the max."""
    if not reference.available():
        pytest.skip(reference.why_unavailable())
    g = torch.Generator().manual_seed(batch_size)
    qs = [torch.nn.functional.normalize(torch.randn(int(n), 128, generator=g), dim=-1).to(dtype)
          for n in torch.randint(3, 33, (21,), generator=g)]
    ps = [torch.nn.functional.normalize(torch.randn(int(n), 128, generator=g), dim=-1).to(dtype)
          for n in torch.randint(1, 300, (45,), generator=g)]
    want = reference.scorer_module().BaseVisualRetrieverProcessor.score_multi_vector(qs, ps, batch_size=batch_size, device="cuda")
    capsys.readouterr()                                           # the reference prints its timings
    got = evdr.BaseVisualRetrieverProcessor.score_multi_vector(qs, ps, batch_size=batch_size, device=str(cuda))
    assert got.device.type == "cpu" and got.dtype == torch.float32 and got.shape == want.shape == (21, 45)
    U.assert_scores_close(got.numpy(), want.numpy(), rel=2e-5 if dtype == torch.float32 else U.REL_TOL)


def _torch_maxsim_restated(Q, P, qmask, pmask, chunk=32):
    """Plain torch fp32 MaxSim with the reference's semantics (evaluator/retrieval.py:166-213),
    written for the tests: one matmul per page chunk, -1e4 fill, max, has-token and query masks, sum."""
    Q, P = Q.float(), P.float()
    qm, pm = qmask.bool(), pmask.bool()
    outs = []
    for s in range(0, P.shape[0], chunk):
        Pc, mc = P[s:s + chunk], pm[s:s + chunk]
        sim = torch.matmul(Q[:, None], Pc[None].transpose(-1, -2))          # (q, c, lq, lp)
        sim = torch.where(mc[None, :, None, :], sim, sim.new_tensor(-1e4))
        mx = sim.max(dim=-1).values * mc.any(1)[None, :, None] * qm[:, None, :]
        outs.append(mx.sum(-1))
    return torch.cat(outs, 1)


# ViDoRe splits in the order of utils/mapping.py:2-91: (pages, queries); query counts are synthetic
# choices; tatdqa's 1663 queries is the largest case.
VIDORE = [("arxivqa", 500, 500), ("docvqa", 500, 500), ("infovqa", 500, 500), ("tabfquad", 70, 280),
          ("tatdqa", 277, 1663), ("shift", 1000, 100), ("ai", 1000, 100), ("energy", 1000, 100),
          ("gov", 1000, 100), ("health", 1000, 100)]


@pytest.mark.parametrize("name,npg,nq", VIDORE)
@pytest.mark.parametrize("lp", [1030, 64])
def test_config2_vidore_shapes(evdr, cuda, name, npg, nq, lp):
    """config 2: all ten splits, full 1030-token pages vs compressed 64-token pages, masked path."""
    seed = 1000 + 2 * [v[0] for v in VIDORE].index(name) + (1 if lp == 64 else 0)
    P, pm = evdr.synth.make_pages(npg, lp, seed=seed, device=cuda, n_holes=6 if lp > 64 else 2,
                                  empty_frac=0.01, min_len=lp // 2)
    Q, qm, rel = evdr.synth.make_queries(P, pm, nq, 32, seed=seed + 1, sigma=3.0 if lp > 64 else 1.5,
                                         min_len=8, drop=0.1)
    got = evdr.score_multi_vector_masked(Q, P, qm, pm, chunk_p=64)
    # the full torch comparison materialises (q, c, 32, lp): check a slab of queries against it ...
    qs = slice(0, min(nq, 48))
    ref = torch_maxsim(Q[qs], P, qm[qs], pm)
    U.assert_scores_close(got[qs].cpu().numpy(), ref.cpu().numpy())
    # ... and everything against the exact fp32 kernels
    exact = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    U.assert_scores_close(got.cpu().numpy(), exact.cpu().numpy())
    # retrieval quality is identical: same top-1 wherever the margin exceeds the tolerance
    top2 = exact.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 2e-3 * top2[:, 0].abs().clamp_min(1)
    assert torch.equal(got.argmax(1)[clear], exact.argmax(1)[clear])


def listwise_distillation_loss(score_s, score_t, k=10, temperature=1.0):
    """The reference's own criterion.py:114-142 when staged, else restated for the test."""
    if reference.available():
        return reference.criterion_module().listwise_distillation_loss(score_s, score_t, k=k, temperature=temperature)
    score_t = score_t.detach()
    log_ps = F.log_softmax(score_s / temperature, dim=1)
    pt = F.softmax(score_t / temperature, dim=1)
    _, idx = torch.topk(pt, k, dim=1)
    return -(pt.gather(1, idx) * log_ps.gather(1, idx)).sum(1).mean() * temperature ** 2


@pytest.mark.parametrize("npg,ls", [(500, 206), (1000, 41), (277, 21)])
def test_config5_training_step(evdr, cuda, npg, ls):
    """One optimisation step of mainv2_iter_liscore.py (:282-310): teacher scores (no grad, tcgen05),
    student scores through l2_normalize(P*mask) with grad (exact kernels), listwise + MSE loss,
    backward into the student parameter -- against the same step in plain torch."""
    B = 32
    Pt, pmt = evdr.synth.make_pages(npg, 1030, seed=1, device=cuda, empty_frac=0.0)
    Q, qm, rel = evdr.synth.make_queries(Pt, pmt, B, 32, seed=2, sigma=3.0, min_len=10, drop=0.05)
    g = torch.Generator(device=cuda).manual_seed(3)
    Pbar = torch.randn(npg, ls, 128, generator=g, device=cuda)
    pms = torch.rand(npg, ls, generator=g, device=cuda) > 0.1
    pms[:, 0] = True

    def step(score_fn, l2n):
        param = Pbar.clone().requires_grad_(True)
        Psb = l2n(param, pms)
        with torch.no_grad():
            sc_t = score_fn(Q, Pt, qm, pmt)
        sc_s = score_fn(Q, Psb, qm, pms)
        loss = listwise_distillation_loss(sc_s, sc_t, k=10, temperature=1.0) + F.mse_loss(sc_s, sc_t)
        loss.backward()
        return sc_t, sc_s.detach(), loss.detach(), param.grad

    def l2n_torch(x, m):
        x = x * m.unsqueeze(-1)
        return x / (x.norm(dim=-1, keepdim=True) + 1e-12)

    t_ref, s_ref, loss_ref, g_ref = step(torch_maxsim, l2n_torch)
    t_got, s_got, loss_got, g_got = step(lambda *a: evdr.score_multi_vector_masked(*a, 64),
                                         lambda x, m: evdr.l2_normalize_masked(x, m))
    U.assert_scores_close(t_got.cpu().numpy(), t_ref.cpu().numpy())
    # student scores: bf16x3 tensor-core search + fp32 rescoring of the winners
    np.testing.assert_allclose(s_got.cpu().numpy(), s_ref.cpu().numpy(), rtol=2e-5, atol=2e-5)
    assert abs(loss_got.item() - loss_ref.item()) <= 2e-3 * max(1.0, abs(loss_ref.item()))
    # masked student tokens receive exactly zero gradient (the invariant mainv1.py:74-87 prints)
    assert g_got[~pms].abs().max() == 0
    # gradients agree except where a near-tie picks another argmax token
    num = (g_got - g_ref).norm()
    den = g_ref.norm()
    assert (num / den).item() < 2e-2, (num / den).item()
    close = torch.isclose(g_got, g_ref, rtol=1e-3, atol=1e-6).float().mean().item()
    assert close > 0.995, close


def test_config1_docvqa_end_to_end(evdr, cuda):
    """config 1: 500 queries x 500 pages, MaxSim + nDCG@5 through the evaluator on device vs the oracle's
    trec restatement applied to plain-torch scores."""
    from oracle import trec_oracle as to
    P, pm = evdr.synth.make_pages(500, 1030, seed=11, device=cuda, n_holes=6, empty_frac=0.01)
    Q, qm, rel = evdr.synth.make_queries(P, pm, 500, 32, seed=12, sigma=3.5, min_len=8, drop=0.1)
    got = evdr.score_multi_vector_masked(Q, P, qm, pm, 64)
    ref = torch.cat([torch_maxsim(Q[i:i + 50], P, qm[i:i + 50], pm) for i in range(0, 500, 50)])
    U.assert_scores_close(got.cpu().numpy(), ref.cpu().numpy())
    docs = [f"doc{c:04d}" for c in range(500)]
    qrels = {f"q{q}": {docs[int(rel[q])]: 1} for q in range(500)}
    refn, gotn = ref.cpu().numpy(), got.cpu().numpy()
    res_ref = {f"q{q}": {docs[c]: float(refn[q, c]) for c in range(500)} for q in range(500)}
    res_got = {f"q{q}": {docs[c]: float(gotn[q, c]) for c in range(500)} for q in range(500)}
    want = to.compute_mteb_metrics(qrels, res_ref)
    have = evdr.CustomRetrievalEvaluator().compute_mteb_metrics(qrels, res_got)
    assert have["NDCG"]["NDCG@5"] == want["NDCG"]["NDCG@5"]
    assert have["Recall"]["Recall@1"] == want["Recall"]["Recall@1"]
    assert 0.3 < want["NDCG"]["NDCG@5"] < 1.0, want["NDCG"]


@pytest.mark.parametrize("route", ["single_pass", "rescore", "composed"])
def test_rescore_path_vs_exact_on_fp32_operands(evdr, cuda, route, monkeypatch):
    """Gradient calls run ONE bf16x3 (near-fp32) tensor-core pass for scores and winners ("single_pass",
    the default), optionally recompute the winners' dot products exactly in fp32 (EVDR_TRAIN_RESCORE=1), or
    -- beyond one evdr_maxsim_train_fwd call -- issue the steps from Python with rescoring in page blocks
    ("composed").  On operands that are NOT bf16-representable every route must equal the exact fp32 path:
    same scores, same winners (up to ties at the 1e-6 level), same gradients."""
    composed = route == "composed"
    if composed:
        import evdr_b200.retrieval as R
        monkeypatch.setattr(R, "_EXACT_WS_BYTES", 1 << 19)
    if route == "rescore":
        monkeypatch.setenv("EVDR_TRAIN_RESCORE", "1")
    g = torch.Generator(device=cuda).manual_seed(5)
    npg, ls, nq = 300, 103, 32
    P = torch.randn(npg, ls, 128, generator=g, device=cuda)
    pm = torch.rand(npg, ls, generator=g, device=cuda) > 0.15
    pm[7] = False                                            # an empty page
    P = evdr.l2_normalize_masked(P, pm)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g, device=cuda), dim=-1)
    qm = torch.rand(nq, 32, generator=g, device=cuda) > 0.1
    gout = torch.randn(nq, npg, generator=g, device=cuda)
    outs, grads = [], []
    for prec in (None, "fp32"):
        Pg = P.clone().requires_grad_(True)
        Qg = Q.clone().requires_grad_(True)
        o = evdr.score_multi_vector_masked(Qg, Pg, qm, pm, 64, precision=prec)
        o.backward(gout)
        outs.append(o.detach())
        grads.append((Pg.grad, Qg.grad))
    np.testing.assert_allclose(outs[0].cpu().numpy(), outs[1].cpu().numpy(), rtol=2e-5, atol=2e-5)
    if route != "single_pass":
        assert (outs[0] <= outs[1] + 1e-5).all(), "rescoring a non-winner can only lower a score"
    assert (outs[0][:, 7] == 0).all() and grads[0][0][7].abs().max() == 0
    assert grads[0][0][~pm].abs().max() == 0, "masked tokens receive exactly zero gradient"
    for a, b in zip(grads[0], grads[1]):
        rel = ((a - b).norm() / b.norm()).item()
        frac = torch.isclose(a, b, rtol=1e-4, atol=1e-6).float().mean().item()
        assert rel < 1e-2 and frac > 0.995, (rel, frac)


@pytest.mark.parametrize("lq", [33, 50, 64, 100])
def test_long_query_gradient_calls_use_tensor_cores(evdr, cuda, lq, monkeypatch):
    """ColPali queries run to 50 tokens (process_queries max_length, evaluator/retrieval.py:63): gradient
    calls with 33..128 query tokens are padded to 32-token groups for the tensor-core search and must equal
    the exact fp32 path in scores and in dP / dQ (which keep the caller's (nq, lq, d) shape)."""
    g = torch.Generator(device=cuda).manual_seed(lq)
    npg, ls, nq = 120, 70, 9
    P = torch.randn(npg, ls, 128, generator=g, device=cuda)
    pm = torch.rand(npg, ls, generator=g, device=cuda) > 0.2
    pm[3] = False
    P = evdr.l2_normalize_masked(P, pm)
    Q = torch.nn.functional.normalize(torch.randn(nq, lq, 128, generator=g, device=cuda), dim=-1)
    qm = torch.rand(nq, lq, generator=g, device=cuda) > 0.15
    gout = torch.randn(nq, npg, generator=g, device=cuda)
    import evdr_b200.retrieval as R
    calls = []
    orig = R._MaxSimRescoreFn.forward
    monkeypatch.setattr(R._MaxSimRescoreFn, "forward", staticmethod(lambda ctx, *a: (calls.append(1), orig(ctx, *a))[1]))
    outs, grads = [], []
    for prec in (None, "fp32"):
        Pg = P.clone().requires_grad_(True)
        Qg = Q.clone().requires_grad_(True)
        o = evdr.score_multi_vector_masked(Qg, Pg, qm, pm, 64, precision=prec)
        o.backward(gout)
        outs.append(o.detach())
        grads.append((Pg.grad, Qg.grad))
    assert len(calls) == 1, "the default precision must take the tensor-core search + rescore path"
    assert grads[0][1].shape == (nq, lq, 128)
    np.testing.assert_allclose(outs[0].cpu().numpy(), outs[1].cpu().numpy(), rtol=2e-5, atol=2e-5)
    assert (outs[0][:, 3] == 0).all()
    assert grads[0][1][~qm].abs().max() == 0, "masked query tokens receive exactly zero gradient"
    for a, b in zip(grads[0], grads[1]):
        rel = ((a - b).norm() / b.norm()).item()
        frac = torch.isclose(a, b, rtol=1e-4, atol=1e-6).float().mean().item()
        assert rel < 1e-2 and frac > 0.995, (rel, frac)


@pytest.mark.parametrize("nq,npg,lp", [(13, 200, 1030), (40, 500, 64), (5, 60, 333)])
def test_bf16x3_high_fidelity_mode(evdr, cuda, nq, npg, lp):
    """precision='bf16x3': K = 384 split operands on the tensor cores reproduce the fp32 scores of
    operands that are not bf16-representable to ~1e-6, where plain bf16 is at ~1e-4."""
    g = torch.Generator(device=cuda).manual_seed(nq)
    P = torch.nn.functional.normalize(torch.randn(npg, lp, 128, generator=g, device=cuda), dim=-1)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g, device=cuda), dim=-1)
    pm = torch.rand(npg, lp, generator=g, device=cuda) > 0.1
    qm = torch.rand(nq, 32, generator=g, device=cuda) > 0.1
    exact = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    x3 = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="bf16x3")
    bf = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="bf16")
    e3 = ((x3 - exact).abs() / exact.abs().clamp_min(1)).max().item()
    e1 = ((bf - exact).abs() / exact.abs().clamp_min(1)).max().item()
    assert e3 < 5e-6, e3
    assert e1 < 1e-3 and e3 < e1
    # fp32 page tensors take the split by default (retrieval.default_precision); a bf16 page tensor plain bf16
    assert torch.equal(evdr.score_multi_vector_masked(Q, P, qm, pm), x3)
    assert torch.equal(evdr.score_multi_vector_masked(Q, P.bfloat16(), qm, pm),
                       evdr.score_multi_vector_masked(Q, P.bfloat16(), qm, pm, precision="bf16"))


def test_hard_token_mining_matches_reference_loop(evdr, cuda):
    """mainv3_iter_liscore_QA_hardtoken.py:398-425 restated in torch (the per-pair loop) vs one launch."""
    P, pm = evdr.synth.make_pages(60, 333, seed=3, device=cuda, n_holes=5, empty_frac=0.0, min_len=40)
    Q, qm, rel = evdr.synth.make_queries(P, pm, 32, 32, seed=4, sigma=2.0, min_len=6, drop=0.1)
    g = torch.Generator(device=cuda).manual_seed(5)
    doc_idx = torch.randint(0, 60, (32, 4), generator=g, device=cuda)
    got = evdr.hard_token_argmax(Q, qm, P, pm, doc_idx)
    Pf, Qf = P.float(), Q.float()
    for qi in range(32):
        qt = Qf[qi][qm[qi]]
        for a in range(4):
            dmask = pm[doc_idx[qi, a]]
            sim = qt @ Pf[doc_idx[qi, a]].T
            sim[:, ~dmask] = float("-inf")
            want = torch.argmax(sim.max(dim=0).values)
            if got[qi, a] != want:   # only an exact-tie / last-bit difference may pick another token
                col = sim.max(dim=0).values
                assert abs(col[got[qi, a]] - col[want]) < 1e-6
    toks = evdr.hard_tokens(Q, qm, P, pm, doc_idx)
    assert toks.shape == (128, 1, 128)
    assert torch.equal(toks[5, 0], P[doc_idx.reshape(-1)[5], got.reshape(-1)[5]])


def test_teacher_score_table_rows_equal_per_batch_calls(evdr, cuda):
    """scoring every pseudo-query against the frozen teacher once and gathering rows
    by dataset index gives the teacher scores of the per-step call of mainv2_iter_listwise.py:283-284
    (same products and maxima; the 32-token sum may be ordered differently between kernel flavours)."""
    P, pm = evdr.synth.make_pages(150, 333, seed=21, device=cuda, n_holes=4, empty_frac=0.02, min_len=50)
    Q, qm, _ = evdr.synth.make_queries(P, pm, 300, 32, seed=22, sigma=2.0, min_len=6, drop=0.1)
    table = evdr.TeacherScoreTable(Q, qm, P, pm, query_block=128)
    assert len(table) == 300
    idx = torch.tensor([5, 299, 0, 5, 128, 127], device=cuda)
    with torch.no_grad():
        want = evdr.score_multi_vector_masked(Q[idx], P, qm[idx], pm, 32)
    torch.testing.assert_close(table[idx], want, rtol=0, atol=2e-5)
    torch.testing.assert_close(table[[7]], evdr.score_multi_vector_masked(Q[7:8], P, qm[7:8], pm), rtol=0, atol=2e-5)


def test_training_step_captured_in_cuda_graph(evdr, cuda):
    """A whole distillation step (masked l2-normalise, teacher + student scoring, fused losses, backward,
    AdamW) captured once with torch.cuda.graph and replayed on new query batches: every libevdr launch goes
    to the capturing stream, scratch comes from stream-ordered allocations, TMA descriptors are encoded at
    capture time against static tensors.  Losses and the trained student are bit-identical to the eager loop."""
    B, npg, ls, steps = 8, 120, 41, 12
    Pt, pmt = evdr.synth.make_pages(npg, 333, seed=1, device=cuda, n_holes=3, empty_frac=0.0, min_len=40)
    Qall, qmall, _ = evdr.synth.make_queries(Pt, pmt, 64, 32, seed=2, sigma=2.0, min_len=6, drop=0.1)
    Qall, Ptf = Qall.float(), Pt.float()
    init = torch.randn(npg, ls, 128, generator=torch.Generator(device=cuda).manual_seed(3), device=cuda)
    pms = torch.rand(npg, ls, generator=torch.Generator(device=cuda).manual_seed(4), device=cuda) > 0.1

    def make():
        p = init.clone().requires_grad_(True)
        return p, torch.optim.AdamW([p], lr=1e-2, capturable=True)

    def step(param, opt, Qb, qmb):
        Psb = evdr.l2_normalize_masked(param, pms)
        with torch.no_grad():
            sc_t = evdr.score_multi_vector_masked(Qb, Ptf, qmb, pmt, 64)
        sc_s = evdr.score_multi_vector_masked(Qb, Psb, qmb, pms, 64)
        ll, lm = evdr.liscore_losses(sc_s, sc_t, k=10, temperature=1.0)
        loss = ll + 0.5 * lm
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        return loss.detach()

    def reset(param, opt):
        param.data.copy_(init)
        for st in opt.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()

    batches = [(Qall[(i * B) % 64:(i * B) % 64 + B], qmall[(i * B) % 64:(i * B) % 64 + B]) for i in range(steps)]
    pe, oe = make()
    eager = [step(pe, oe, Qb, qmb).clone() for Qb, qmb in batches]

    pg, og = make()
    Qs, qms = Qall[:B].clone(), qmall[:B].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):                                     # warm-up: optimizer state, caches
            step(pg, og, Qs, qms)
    torch.cuda.current_stream().wait_stream(side)
    reset(pg, og)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_static = step(pg, og, Qs, qms)
    reset(pg, og)
    replayed = []
    for Qb, qmb in batches:
        Qs.copy_(Qb)
        qms.copy_(qmb)
        graph.replay()
        replayed.append(loss_static.clone())
    assert torch.equal(torch.stack(eager), torch.stack(replayed))
    assert torch.equal(pe.detach(), pg.detach())
    assert bool(torch.isfinite(torch.stack(eager)).all()) and not torch.equal(pe.detach(), init), "the student was trained"


def test_reference_eval_loop_reads_scores_from_host_mirror(evdr, cuda):
    """The reference builds its results dict with float(scores[qi, di].item()) for every pair
    (mainv2_iter_listwise.py:312-318).  No-grad results are ordinary CUDA tensors whose scalar reads come from
    a host mirror (one copy instead of Q*P device synchronisations); everything else behaves like a Tensor."""
    import time
    P, pm = evdr.synth.make_pages(300, 64, seed=31, device=cuda, n_holes=2, empty_frac=0.01, min_len=20)
    Q, qm, _ = evdr.synth.make_queries(P, pm, 200, 32, seed=32, sigma=2.0, min_len=6, drop=0.1)
    with torch.no_grad():
        scores = evdr.score_multi_vector_masked(Q, P, qm, pm, chunk_p=64)
    assert isinstance(scores, torch.Tensor) and scores.is_cuda and scores.dtype == torch.float32
    plain = scores + 0                                     # ops return plain tensors
    assert type(plain) is torch.Tensor and type(scores[3]) is torch.Tensor and type(scores[:, 1:5]) is torch.Tensor
    t0 = time.perf_counter()
    results = {str(qi): {str(di): float(scores[qi, di].item()) for di in range(scores.shape[1])}
               for qi in range(scores.shape[0])}
    dt = time.perf_counter() - t0
    host = plain.cpu()
    assert all(results[str(q)][str(c)] == float(host[q, c]) for q in (0, 7, 199) for c in (0, 150, 299))
    assert dt < 3.0, f"{dt:.2f} s for 60 000 scalar reads: the mirror is not being used"
    scores.mul_(2.0)                                       # an in-place write invalidates the mirror
    assert float(scores[7, 150].item()) == 2.0 * float(host[7, 150])
    # results that carry a gradient are untouched
    Pg = P.float().requires_grad_(True)
    assert type(evdr.score_multi_vector_masked(Q[:4], Pg, qm[:4], pm)) is torch.Tensor


def test_gradient_call_over_more_than_65535_pages(evdr, cuda):
    """A single query against 70 000 short student pages: beyond the page count one rescoring launch covers,
    so the composed path (page blocks) must take over transparently."""
    g = torch.Generator(device=cuda).manual_seed(9)
    npg, ls = 70_000, 5
    P = torch.nn.functional.normalize(torch.randn(npg, ls, 128, generator=g, device=cuda), dim=-1)
    pm = torch.rand(npg, ls, generator=g, device=cuda) > 0.2
    Q = torch.nn.functional.normalize(torch.randn(1, 7, 128, generator=g, device=cuda), dim=-1)
    qm = torch.ones(1, 7, dtype=torch.bool, device=cuda)
    Pg = P.clone().requires_grad_(True)
    out = evdr.score_multi_vector_masked(Q, Pg, qm, pm)
    out.sum().backward()
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-5)
    assert Pg.grad.shape == P.shape and Pg.grad[~pm].abs().max() == 0
