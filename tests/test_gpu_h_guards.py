"""Out-of-bounds WRITE detection without a sanitizer (compute-sanitizer is closed on this GPU pool): every
output of the library is placed inside a larger buffer pre-filled with a sentinel, with guard cells before,
after and -- for strided outputs -- between the rows; after the call every guard cell must be untouched and
every output cell written."""
import ctypes as C

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("plain_bf16")]

SENT_F = -12345.678
SENT_I = -987654321


def _guarded(shape, dtype, cuda, pad_rows=2, pad_cols=24):
    """(big, view): view has `shape` and sits inside big with guard rows / columns all around."""
    rows, cols = shape
    sent = SENT_F if dtype.is_floating_point else SENT_I
    big = torch.full((rows + 2 * pad_rows, cols + 2 * pad_cols), sent, dtype=dtype, device=cuda)
    view = big[pad_rows:pad_rows + rows, pad_cols:pad_cols + cols]
    return big, view, sent


def _check(big, view, sent, what):
    mask = torch.ones_like(big, dtype=torch.bool)
    pr = (big.shape[0] - view.shape[0]) // 2
    pc = (big.shape[1] - view.shape[1]) // 2
    mask[pr:pr + view.shape[0], pc:pc + view.shape[1]] = False
    assert bool((big[mask] == sent).all()), f"{what}: a guard cell was overwritten"
    assert bool((view != sent).all()), f"{what}: an output cell was never written"


@pytest.mark.parametrize("nq,npg,lp", [(1, 37, 1030), (4, 300, 64), (7, 200, 1030), (8, 129, 333), (12, 90, 1030),
                                       (33, 257, 16), (64, 150, 1030), (100, 999, 64), (130, 61, 2100)])
def test_scores_stay_inside_their_block(evdr, cuda, nq, npg, lp):
    P, pm = evdr.synth.make_pages(npg, lp, seed=nq + lp, device=cuda, n_holes=min(3, lp - 1), empty_frac=0.02,
                                  min_len=max(1, lp // 4))
    Q, qm, _ = evdr.synth.make_queries(P, pm, nq, 32, seed=nq, sigma=2.0, min_len=4, drop=0.1)
    pk = evdr.PackedPages.from_padded(P, pm, varlen=True)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    big, out, sent = _guarded((nq, npg), torch.float32, cuda)
    pk.score_packed(Qb, qm32, out=out)
    _check(big, out, sent, "evdr_maxsim_packed_fwd")
    ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    assert torch.allclose(out, ref, rtol=0, atol=2e-2)


@pytest.mark.parametrize("nq,n,k", [(1, 100000, 10), (3, 70, 32), (64, 5000, 10), (5, 9, 100), (200, 333, 17)])
def test_topk_outputs_stay_inside_their_block(evdr, cuda, nq, n, k):
    lib = evdr._lib.load()
    h = evdr._lib.handle(cuda)
    st = evdr._lib.stream_ptr(cuda)
    bigs, s, sent = _guarded((nq, n), torch.float32, cuda)
    s.copy_(torch.randn(nq, n, device=cuda))
    big_v, tv, sv = _guarded((nq, k), torch.float32, cuda)
    big_i, ti, si = _guarded((nq, k), torch.int64, cuda)
    # the library writes dense [nq, k] outputs: give it dense scratch inside a guarded 1-D buffer instead
    flat_v = torch.full((nq * k + 64,), SENT_F, dtype=torch.float32, device=cuda)
    flat_i = torch.full((nq * k + 64,), SENT_I, dtype=torch.int64, device=cuda)
    evdr._lib.check(lib.evdr_topk(h, s.data_ptr(), s.stride(0), nq, n, k, None, 0, flat_v[32:].data_ptr(),
                                  flat_i[32:].data_ptr(), st), "evdr_topk")
    assert bool((flat_v[:32] == SENT_F).all() and (flat_v[32 + nq * k:] == SENT_F).all())
    assert bool((flat_i[:32] == SENT_I).all() and (flat_i[32 + nq * k:] == SENT_I).all())
    assert bool((bigs[:2] == sent).all() and (bigs[-2:] == sent).all()), "the score matrix is read-only"
    want = torch.topk(s, min(k, n), dim=1).values
    assert torch.equal(flat_v[32:32 + nq * k].view(nq, k)[:, : min(k, n)], want)


@pytest.mark.parametrize("rescore", [0, 1])
@pytest.mark.parametrize("nq,npg,ls,lq", [(32, 100, 41, 32), (5, 33, 206, 32), (9, 64, 21, 64), (1, 7, 1, 1)])
def test_training_forward_and_backward_stay_inside_their_buffers(evdr, cuda, nq, npg, ls, lq, rescore):
    lib = evdr._lib.load()
    h = evdr._lib.handle(cuda)
    st = evdr._lib.stream_ptr(cuda)
    g = torch.Generator(device=cuda).manual_seed(nq * 7 + ls)
    P = torch.nn.functional.normalize(torch.randn(npg, ls, 128, generator=g, device=cuda), dim=-1)
    Q = torch.nn.functional.normalize(torch.randn(nq, lq, 128, generator=g, device=cuda), dim=-1)
    pm = (torch.rand(npg, ls, generator=g, device=cuda) > 0.2).view(torch.uint8)
    qm = (torch.rand(nq, lq, generator=g, device=cuda) > 0.1).view(torch.uint8)
    R = nq * lq
    pad = 64
    am = torch.full((npg * R + 2 * pad,), SENT_I, dtype=torch.int32, device=cuda)
    out = torch.full((nq * npg + 2 * pad,), SENT_F, dtype=torch.float32, device=cuda)
    evdr._lib.check(lib.evdr_maxsim_train_fwd(h, Q.data_ptr(), 0, P.data_ptr(), 0, qm.data_ptr(), pm.data_ptr(), nq, npg,
                                              lq, ls, 128, rescore, am[pad:].data_ptr(), out[pad:].data_ptr(), st),
                    "evdr_maxsim_train_fwd")
    for buf, sent, n in ((am, SENT_I, npg * R), (out, SENT_F, nq * npg)):
        assert bool((buf[:pad] == sent).all() and (buf[pad + n:] == sent).all()), "guard overwritten"
        assert bool((buf[pad:pad + n] != sent).all()), "output cell never written"
    assert bool((am[pad:pad + npg * R] >= 0).all() and (am[pad:pad + npg * R] < ls).all()), "winner index out of range"
    dP = torch.full((npg * ls * 128 + 2 * pad,), SENT_F, dtype=torch.float32, device=cuda)
    dQ = torch.full((nq * lq * 128 + 2 * pad,), SENT_F, dtype=torch.float32, device=cuda)
    gout = torch.randn(nq, npg, generator=g, device=cuda)
    evdr._lib.check(lib.evdr_maxsim_bwd(h, gout.data_ptr(), Q.data_ptr(), 0, P.data_ptr(), 0, qm.data_ptr(),
                                        pm.data_ptr(), am[pad:].data_ptr(), nq, npg, lq, ls, 128, dP[pad:].data_ptr(),
                                        dQ[pad:].data_ptr(), st), "evdr_maxsim_bwd")
    for buf, n in ((dP, npg * ls * 128), (dQ, nq * lq * 128)):
        assert bool((buf[:pad] == SENT_F).all() and (buf[pad + n:] == SENT_F).all()), "guard overwritten"
        assert bool((buf[pad:pad + n] != SENT_F).all()), "gradient cell never written"


@pytest.mark.parametrize("npg,lp,varlen,kb", [(50, 1030, 0, 1), (200, 64, 1, 1), (33, 333, 1, 3), (7, 1, 0, 3)])
def test_packed_stream_stays_inside_its_buffer(evdr, cuda, npg, lp, varlen, kb):
    lib = evdr._lib.load()
    h = evdr._lib.handle(cuda)
    st = evdr._lib.stream_ptr(cuda)
    g = torch.Generator(device=cuda).manual_seed(npg + lp)
    P = torch.randn(npg, lp, 128, generator=g, device=cuda)
    pm = (torch.rand(npg, lp, generator=g, device=cuda) > 0.3).view(torch.uint8)
    off = torch.full((npg + 1 + 16,), SENT_I, dtype=torch.int32, device=cuda)
    evdr._lib.check(lib.evdr_pack_plan(h, pm.data_ptr(), npg, lp, varlen, off[8:].data_ptr(), st), "evdr_pack_plan")
    assert bool((off[:8] == SENT_I).all() and (off[8 + npg + 1:] == SENT_I).all())
    total = int(off[8 + npg].item())
    assert total >= npg
    width = 128 * evdr._lib.row_blocks(kb)
    pad = 4 * width
    tok = torch.full((total * 16 * width + 2 * pad,), 7.0, dtype=torch.bfloat16, device=cuda)
    evdr._lib.check(lib.evdr_pack_pages(h, P.data_ptr(), 0, pm.data_ptr(), off[8:].data_ptr(), npg, lp, 128, 1, kb,
                                        tok[pad:].data_ptr(), st), "evdr_pack_pages")
    assert bool((tok[:pad] == 7.0).all() and (tok[pad + total * 16 * width:] == 7.0).all()), "guard overwritten"
    body = tok[pad:pad + total * 16 * width].float()
    assert bool((body.abs() <= 1.0001).all()), "normalised tokens (or zero fillers) only: a 7.0 survived inside"
