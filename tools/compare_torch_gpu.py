"""The reference's own eager path on the same B200 ("the bar is the reference's own torch
path") beside the drop-in, for config 1 (500 x 500 eval), config-2 shapes and a config-5 training step.
The eager path is the UNMODIFIED evaluator/retrieval.py:166-213 staged under baseline/_ref (fp32,
chunk_p=64); the drop-in is timed with its default precision for fp32 tensors (bf16x3 split) and with
EVDR_PRECISION=bf16."""
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr
import os
from baseline import ref as reference
eager = reference.scorer_module().score_multi_vector_masked          # the reference's own function

dev = torch.device("cuda", 0)


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for nq, npg, lp in [(500, 500, 1030), (32, 1000, 1030), (1663, 277, 1030), (500, 500, 64)]:
    P, pm = evdr.synth.make_pages(npg, lp, seed=1, device=dev)
    Q, qm, rel = evdr.synth.make_queries(P, pm, nq, 32, seed=2, sigma=3.0)
    Pf, Qf = P.float(), Q.float()
    qb = 64  # the (q, 64 pages, 32, lp) fp32 intermediate of the eager path is 4.2 GB at 500 queries: batch queries
    def eager_all():
        return torch.cat([eager(Qf[i:i + qb], Pf, qm[i:i + qb], pm, 64) for i in range(0, nq, qb)])
    t_e = timeit(eager_all, n=3, warm=1)
    t_o = timeit(lambda: evdr.score_multi_vector_masked(Qf, Pf, qm, pm, 64), n=20, warm=3)
    t_b = timeit(lambda: evdr.score_multi_vector_masked(Qf, Pf, qm, pm, 64, precision="bf16"), n=20, warm=3)
    evdr.retrieval.clear_pack_cache()
    t_o_cold = timeit(lambda: (evdr.retrieval.clear_pack_cache(), evdr.score_multi_vector_masked(Qf, Pf, qm, pm, 64)), n=20, warm=3)
    err = ((evdr.score_multi_vector_masked(Qf, Pf, qm, pm) - eager_all()).abs() / eager_all().abs().clamp_min(1)).max().item()
    print(f"scoring {nq}x{npg} Lp={lp}: reference torch eager on B200 {t_e:9.3f} ms | drop-in (fp32 pages -> bf16x3) {t_o:7.3f} ms "
          f"(packed-cache hit) {t_o_cold:7.3f} ms (pack every call) | x{t_e/t_o:7.1f} / x{t_e/t_o_cold:6.1f} | max rel err {err:.1e} "
          f"| precision=bf16: {t_b:7.3f} ms x{t_e/t_b:7.1f}", flush=True)

# training step, config 5
B, npg, ls = 32, 1000, 41
Pt, pmt = evdr.synth.make_pages(npg, 1030, seed=1, device=dev, empty_frac=0.0)
Q, qm, rel = evdr.synth.make_queries(Pt, pmt, B, 32, seed=2, sigma=3.0)
Ptf, Qf = Pt.float(), Q.float()
param = torch.randn(npg, ls, 128, device=dev, requires_grad=True)
pms = torch.ones(npg, ls, dtype=torch.bool, device=dev)
opt = torch.optim.AdamW([param], lr=1e-3)


def loss_fn(sc_s, sc_t):
    log_ps = F.log_softmax(sc_s, dim=1)
    pt = F.softmax(sc_t, dim=1)
    _, idx = torch.topk(pt, 10, dim=1)
    return -(pt.gather(1, idx) * log_ps.gather(1, idx)).sum(1).mean() + F.mse_loss(sc_s, sc_t)


def step(score, l2n):
    Psb = l2n(param)
    with torch.no_grad():
        sc_t = score(Qf, Ptf, qm, pmt)
    sc_s = score(Qf, Psb, qm, pms)
    loss = loss_fn(sc_s, sc_t)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


l2_t = lambda x: (x * pms.unsqueeze(-1)) / ((x * pms.unsqueeze(-1)).norm(dim=-1, keepdim=True) + 1e-12)
t_e = timeit(lambda: step(lambda *a: eager(*a, 64), l2_t), n=5, warm=2)
t_o = timeit(lambda: step(lambda *a: evdr.score_multi_vector_masked(*a, 64), lambda x: evdr.l2_normalize_masked(x, pms)), n=20, warm=3)
os.environ["EVDR_PRECISION"] = "bf16"
t_b = timeit(lambda: step(lambda *a: evdr.score_multi_vector_masked(*a, 64), lambda x: evdr.l2_normalize_masked(x, pms)), n=20, warm=3)
print(f"training step B={B} N={npg} Lt=1030 Ls={ls}: reference torch eager on B200 {t_e:9.3f} ms | drop-in {t_o:7.3f} ms | x{t_e/t_o:.1f} "
      f"| teacher scored as plain bf16 (EVDR_PRECISION=bf16): {t_b:7.3f} ms x{t_e/t_b:.1f}", flush=True)
