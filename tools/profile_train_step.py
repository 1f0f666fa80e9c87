"""Kernel-level breakdown of one config-5 training step through the drop-in (torch profiler)."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

dev = torch.device("cuda", 0)
B, npg, ls = 32, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 41
Pt, pmt = evdr.synth.make_pages(npg, 1030, seed=1, device=dev, empty_frac=0.0)
Q, qm, rel = evdr.synth.make_queries(Pt, pmt, B, 32, seed=2, sigma=3.0)
Ptf, Qf = Pt.float(), Q.float()
param = torch.randn(npg, ls, 128, device=dev, requires_grad=True)
pms = torch.ones(npg, ls, dtype=torch.bool, device=dev)
opt = torch.optim.AdamW([param], lr=1e-3)


def step():
    Psb = evdr.l2_normalize_masked(param, pms)
    with torch.no_grad():
        sc_t = evdr.score_multi_vector_masked(Qf, Ptf, qm, pmt, 64)
    sc_s = evdr.score_multi_vector_masked(Qf, Psb, qm, pms, 64)
    log_ps = F.log_softmax(sc_s, dim=1)
    pt = F.softmax(sc_t, dim=1)
    _, idx = torch.topk(pt, 10, dim=1)
    loss = -(pt.gather(1, idx) * log_ps.gather(1, idx)).sum(1).mean() + F.mse_loss(sc_s, sc_t)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
