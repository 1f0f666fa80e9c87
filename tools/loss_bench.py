"""Fused score-matrix consumers (losses.cu) vs the same formulas as torch eager ops on the same GPU
(CUDA events, 50 iterations after warm-up).  python tools/loss_bench.py"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

dev = torch.device("cuda", 0)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def torch_liscore(s, t, k, T):
    lps, pt = F.log_softmax(s / T, dim=1), F.softmax(t / T, dim=1)
    _, idx = torch.topk(pt, k, dim=1)
    return -(pt.gather(1, idx) * lps.gather(1, idx)).sum(1).mean() * T * T, F.mse_loss(s, t)


def torch_pairwise(s, t):
    ds, dt = s.unsqueeze(2) - s.unsqueeze(1), t.unsqueeze(2) - t.unsqueeze(1)
    return F.binary_cross_entropy_with_logits(ds, torch.sigmoid(dt))


def torch_rank_gap(sc_t, sc_s, k, a):
    rank_t = torch.argsort(torch.argsort(sc_t, dim=-1, descending=True), dim=-1)
    rank_s = torch.argsort(torch.argsort(sc_s, dim=-1, descending=True), dim=-1)
    gap = rank_t.float() - rank_s.float()
    G = gap.abs().sum(dim=0)
    idx = torch.topk(sc_t, k=k, dim=-1).indices
    pos = torch.topk(gap.gather(1, idx).abs(), k=a, dim=1).indices
    return G, idx.gather(1, pos)


for B, N in ((32, 500), (32, 1000), (300, 277)):
    t = 12 + 3 * torch.randn(B, N, device=dev)
    s0 = t + 1.5 * torch.randn(B, N, device=dev)

    def step(fn):
        def run():
            s = s0.clone().requires_grad_(True)
            out = fn(s)
            (out[0] + out[1] if isinstance(out, tuple) else out).backward()
        return run

    a = timed(step(lambda s: evdr.liscore_losses(s, t, 10, 1.0)))
    b = timed(step(lambda s: torch_liscore(s, t, 10, 1.0)))
    c = timed(step(lambda s: evdr.pairwise_distillation_loss(s, t)))
    d = timed(step(lambda s: torch_pairwise(s, t)))
    e = timed(lambda: evdr.rank_gap(t, s0, 10, 4))
    f = timed(lambda: torch_rank_gap(t, s0, 10, 4))
    print(f"B={B} N={N}: liscore fwd+bwd {a:7.1f} us (torch eager {b:7.1f})   pairwise fwd+bwd {c:7.1f} us "
          f"(torch eager {d:7.1f})   rank-gap {e:7.1f} us (torch eager {f:7.1f})", flush=True)
