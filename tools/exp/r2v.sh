mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2v_sp4.log 2>&1
import os, sys, time, torch
sys.path.insert(0, '.')
import evdr_b200 as evdr
dev = torch.device('cuda', 0)
def run(pages, L, nq, mode, secs=1.5):
    if mode is None: os.environ.pop('EVDR_TC_MODE', None)
    else: os.environ['EVDR_TC_MODE'] = mode
    corpus = evdr.synth.make_uniform_corpus(pages, L, device=dev, seed=1)
    pk = evdr.PackedPages.from_uniform_tokens(corpus)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=dev), dim=-1)
    qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    out = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    os.environ['EVDR_TC_MODE'] = '1cta'
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
        if n % 32 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gbs = pages * L * 256 / (ms * 1e-3) / 1e9
    print(f"L={L:5d} pages={pages:8d} nq={nq} mode={str(mode):5s} scan {ms:8.3f} ms = {gbs:7.1f} GB/s of page bytes  max|diff| {err:.1e}", flush=True)
    del corpus, pk, out
for L, pages in ((16, 6000000), (64, 1600000), (128, 800000), (208, 500000)):
    for nq in (1, 2, 4):
        for mode in ('1cta', 'sp4'):
            run(pages, L, nq, mode)
PY
cat gpurun_out/r2v_sp4.log
