"""Where a top-10 search spends its time: scan alone, top-k alone, both (CUDA events, 20 iterations).
python tools/search_bench.py [pages] [nq ...]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

pages = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nqs = [int(x) for x in sys.argv[2:]] or [1, 64]
dev = torch.device("cuda", 0)
corpus = evdr.synth.make_uniform_corpus(pages, 1040, device=dev, seed=1)
pk = evdr.PackedPages.from_uniform_tokens(corpus)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


for nq in nqs:
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=dev), dim=-1)
    qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    s = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    a, ah = timed(lambda: pk.score_packed(Qb, qm32, out=s))
    b, bh = timed(lambda: evdr.top_k(s, 10))
    c, ch = timed(lambda: evdr.top_k(pk.score_packed(Qb, qm32, out=s), 10))
    d, dh = timed(lambda: pk.search(Q, qm, 10))
    # in-situ split: events between the scan and the top-k of the same iteration
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(20)]
    for e in evs:
        e[0].record()
        pk.score_packed(Qb, qm32, out=s)
        e[1].record()
        evdr.top_k(s, 10)
        e[2].record()
    torch.cuda.synchronize()
    sc = sum(e[0].elapsed_time(e[1]) for e in evs[2:]) / 18
    tk = sum(e[1].elapsed_time(e[2]) for e in evs[2:]) / 18
    print(f"nq={nq:4d} in situ: scan {sc:.3f} ms, top-10 {tk:.3f} ms")
    print(f"nq={nq:4d} pages={pages}: scan {a:8.3f} ms (host {ah:.3f})  top-10 alone {b:8.3f} ms (host {bh:.3f})  "
          f"scan+top-10 {c:8.3f} ms (host {ch:.3f})  PackedPages.search {d:8.3f} ms (host {dh:.3f})", flush=True)
