"""A few launches of the tcgen05 MaxSim kernel on a uniform corpus of L-slot pages (default 1040), for ncu captures.
    [EVDR_TC_MODE=...] python tools/tc_launch.py [pages] [nq] [launches] [L]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

pages = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = int(sys.argv[4]) if len(sys.argv) > 4 else 1040
dev = torch.device("cuda", 0)
corpus = evdr.synth.make_uniform_corpus(pages, L, device=dev, seed=1)
pk = evdr.PackedPages.from_uniform_tokens(corpus)
Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=dev), dim=-1)
qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
Qb, qm32 = evdr.pack_queries(Q, qm)
out = torch.empty((nq, pages), dtype=torch.float32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    pk.score_packed(Qb, qm32, out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
tf = [nq * pages * 2 * 128 * 32 * (1030 if L == 1040 else L) / (m * 1e-3) / 1e12 for m in ms]
gbs = [pages * L * 256 / (m * 1e-3) / 1e9 for m in ms]
print("ms per launch:", [round(m, 3) for m in ms], "TFLOP/s:", [round(t, 1) for t in tf], "GB/s of page slots:", [round(g) for g in gbs])
