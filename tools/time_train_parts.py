"""Wall-clock split of a config-5 training step (CPU launch overhead vs GPU time)."""
import sys, time
from pathlib import Path
import torch, torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr
dev = torch.device("cuda", 0)
B, npg, ls = 32, 1000, 41
Pt, pmt = evdr.synth.make_pages(npg, 1030, seed=1, device=dev, empty_frac=0.0)
Q, qm, rel = evdr.synth.make_queries(Pt, pmt, B, 32, seed=2, sigma=3.0)
Ptf, Qf = Pt.float(), Q.float()
param = torch.randn(npg, ls, 128, device=dev, requires_grad=True)
pms = torch.ones(npg, ls, dtype=torch.bool, device=dev)
opt = torch.optim.AdamW([param], lr=1e-3)

def timed(fn, n=50, sync=True):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if sync: torch.cuda.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def cpu_only(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return t

with torch.no_grad():
    print("teacher call      : %.3f ms sync | %.3f ms cpu-issue" % (timed(lambda: evdr.score_multi_vector_masked(Qf, Ptf, qm, pmt, 64)), cpu_only(lambda: evdr.score_multi_vector_masked(Qf, Ptf, qm, pmt, 64))))
    print("pack_queries      : %.3f ms cpu-issue" % cpu_only(lambda: evdr.pack_queries(Qf, qm)))
def fwd():
    Psb = evdr.l2_normalize_masked(param, pms)
    return evdr.score_multi_vector_masked(Qf, Psb, qm, pms, 64)
print("student fwd       : %.3f ms sync | %.3f ms cpu-issue" % (timed(fwd), cpu_only(fwd)))
def fb():
    s = fwd(); s.sum().backward(); param.grad = None
print("student fwd+bwd   : %.3f ms sync | %.3f ms cpu-issue" % (timed(fb), cpu_only(fb)))
def full():
    Psb = evdr.l2_normalize_masked(param, pms)
    with torch.no_grad(): sc_t = evdr.score_multi_vector_masked(Qf, Ptf, qm, pmt, 64)
    sc_s = evdr.score_multi_vector_masked(Qf, Psb, qm, pms, 64)
    log_ps = F.log_softmax(sc_s, dim=1); pt = F.softmax(sc_t, dim=1); _, idx = torch.topk(pt, 10, dim=1)
    loss = -(pt.gather(1, idx) * log_ps.gather(1, idx)).sum(1).mean() + F.mse_loss(sc_s, sc_t)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
print("full step         : %.3f ms sync | %.3f ms cpu-issue" % (timed(full), cpu_only(full)))
