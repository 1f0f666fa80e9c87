mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2r_short_pages.log 2>&1
import os, sys, time, torch
sys.path.insert(0, '.')
import evdr_b200 as evdr
dev = torch.device('cuda', 0)
def run(label, pages, L, nq, mode, secs=2.0):
    if mode is None: os.environ.pop('EVDR_TC_MODE', None)
    else: os.environ['EVDR_TC_MODE'] = mode
    corpus = evdr.synth.make_uniform_corpus(pages, L, device=dev, seed=1)
    pk = evdr.PackedPages.from_uniform_tokens(corpus)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=dev), dim=-1)
    qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    out = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    os.environ['EVDR_TC_MODE'] = 'ew8' if L >= 512 else 'sp8'
    ref = pk.score_packed(Qb, qm32).clone()
    if mode is None: os.environ.pop('EVDR_TC_MODE', None)
    else: os.environ['EVDR_TC_MODE'] = mode
    for _ in range(3): pk.score_packed(Qb, qm32, out=out)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    n, t0 = 0, time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        pk.score_packed(Qb, qm32, out=out); n += 1
        if n % 16 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = nq * pages * 2 * 128 * 32 * L / (ms * 1e-3) / 1e12
    print(f"{label:28s} L={L:5d} nq={nq:5d} mode={str(mode):7s} {ms:9.3f} ms {tf:8.1f} TFLOP/s (all {L} slots counted)  max|diff| vs single-CTA {err:.2e}", flush=True)
    del corpus, pk, out
for L, pages in ((64, 1600000), (208, 500000), (16, 3000000)):
    for nq in (64, 1024):
        for mode in ('sp8', 'pairsp'):
            run('short pages', pages, L, nq, mode)
for nq in (5, 8):
    for mode in (None, 'pair', 'ta', '1cta'):
        run('5-8 queries', 100000, 1040, nq, mode)
for nq in (9, 12):
    for mode in (None, 'pair', 'ew8'):
        run('9-12 queries', 100000, 1040, nq, mode)
PY
cat gpurun_out/r2r_short_pages.log
