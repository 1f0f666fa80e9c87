"""Kernel-only timing of the tcgen05 MaxSim kernel on a uniform corpus (CUDA events, L2-exceeding
corpus).  python tools/tc_bench.py [pages] [nq ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

import os
pages = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nqs = [int(x) for x in sys.argv[2:]] or [1, 4, 8, 12, 64, 256, 1024]
dev = torch.device("cuda", 0)
TOK = int(os.environ.get("TC_BENCH_TOKENS", "1040"))     # token slots per page (64 = compressed pages)
VALID = 1030 if TOK == 1040 else TOK
corpus = evdr.synth.make_uniform_corpus(pages, TOK, device=dev, seed=1)
pk = evdr.PackedPages.from_uniform_tokens(corpus)
for nq in nqs:
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=dev), dim=-1)
    qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    out = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    for _ in range(3):
        pk.score_packed(Qb, qm32, out=out)
    torch.cuda.synchronize()
    n = 10 if nq <= 64 else 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        pk.score_packed(Qb, qm32, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    pairs = nq * pages / (ms * 1e-3)
    print(f"nq={nq:5d} pages={pages} x {TOK} tok: {ms:9.3f} ms  {pairs/1e6:9.2f} Mpairs/s  "
          f"{pairs*2*128*32*VALID/1e12:8.1f} TFLOP/s(alg)  {pages*VALID*256/(ms*1e-3)/1e9:8.1f} GB/s(alg)", flush=True)
