mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_tests.log; tail -5 gpurun_out/r2w_tests.log
STRESS_SEED=23 timeout 300 python tools/stress_flavours.py 60 > gpurun_out/r2w_stress.log 2>&1; echo stress rc=$?; tail -2 gpurun_out/r2w_stress.log
python - <<'PY' > gpurun_out/r2w_short.log 2>&1
import os, sys, time, torch
sys.path.insert(0, '.')
import evdr_b200 as evdr
dev = torch.device('cuda', 0)
def run(pages, L, nq, secs=1.2):
    corpus = evdr.synth.make_uniform_corpus(pages, L, device=dev, seed=1)
    pk = evdr.PackedPages.from_uniform_tokens(corpus)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=dev), dim=-1)
    qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    out = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    ss = evdr.ShardedSearch((0, pages), mode="nccl")
    res = []
    for fn in (lambda: pk.score_packed(Qb, qm32, out=out), lambda: ss.search_packed(pk, Qb, qm32, 10)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        n, t0 = 0, time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < secs:
            fn(); n += 1
            if n % 32 == 0: torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / n)
    gbs = pages * L * 256 / (res[0] * 1e-3) / 1e9
    print(f"L={L:5d} pages={pages:8d} nq={nq} default dispatch: scan {res[0]:8.3f} ms = {gbs:7.1f} GB/s of page bytes | fused search top-10 {res[1]:8.3f} ms", flush=True)
    del corpus, pk, out
for L, pages in ((16, 6000000), (64, 1600000), (128, 800000), (208, 500000), (512, 200000), (1040, 100000)):
    for nq in (1, 2, 4):
        run(pages, L, nq)
PY
cat gpurun_out/r2w_short.log
