"""Randomised stress of the kernel flavours at sizes where every CTA processes many work items (dynamic
scheduler, query-group changes, A-tile double buffering, pair hand-offs): drop-in scores vs the exact fp32
kernel, fused search vs torch.topk of the unfused score rows, bf16x3 on fp32 tensors.
    python tools/stress_flavours.py [seconds]        # prints one line per case, exits 1 on the first mismatch
"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
dev = torch.device("cuda", 0)
rng = np.random.default_rng(int(os.environ.get("STRESS_SEED", "7")))
NQ = [1, 2, 3, 4, 5, 8, 9, 12, 13, 16, 20, 28, 32, 36, 48, 64, 100, 108, 128, 200, 256, 300]
LP = [16, 40, 64, 100, 208, 333, 511, 520, 1030, 1040, 1500, 2100]
t_end, n_cases = time.time() + secs, 0
while time.time() < t_end:
    nq, lp = int(rng.choice(NQ)), int(rng.choice(LP))
    tokens_total = int(rng.choice([200_000, 800_000, 3_000_000]))
    npg = max(3, tokens_total // lp)
    varlen, fp32_in = bool(rng.integers(0, 2)), bool(rng.integers(0, 4) == 0)
    seed = int(rng.integers(0, 1 << 30))
    g = torch.Generator(device=dev).manual_seed(seed)
    P = torch.nn.functional.normalize(torch.randn(npg, lp, 128, generator=g, device=dev), dim=-1)
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g, device=dev), dim=-1)
    if not fp32_in:
        P, Q = P.bfloat16(), Q.bfloat16()
    pm = torch.rand(npg, lp, generator=g, device=dev) > 0.15
    lens = torch.randint(max(1, lp // 3), lp + 1, (npg,), generator=g, device=dev)
    pm &= torch.arange(lp, device=dev)[None, :] < lens[:, None]
    pm[torch.rand(npg, generator=g, device=dev) < 0.01] = False
    qm = torch.rand(nq, 32, generator=g, device=dev) > 0.1
    evdr.retrieval.clear_pack_cache()
    got = evdr.score_multi_vector_masked(Q, P, qm, pm, 64)
    # exact fp32 reference on a sub-block (the SIMT kernel is slow): random page slab x up to 24 queries
    c0 = int(rng.integers(0, max(1, npg - 400)))
    cs, qs = slice(c0, min(npg, c0 + 400)), slice(0, min(nq, 24))
    ref = evdr.score_multi_vector_masked(Q[qs], P[cs], qm[qs], pm[cs], precision="fp32")
    err = ((got[qs, cs] - ref).abs() / ref.abs().clamp_min(1)).max().item()
    tol = 2e-5 if fp32_in else 1e-3
    ok = err <= tol
    # a second slab at the END of the corpus (last items, ragged tail)
    cs2 = slice(max(0, npg - 200), npg)
    ref2 = evdr.score_multi_vector_masked(Q[qs], P[cs2], qm[qs], pm[cs2], precision="fp32")
    err2 = ((got[qs, cs2] - ref2).abs() / ref2.abs().clamp_min(1)).max().item()
    ok = ok and err2 <= tol
    # fused search vs torch.topk of the unfused rows of the SAME packed stream
    srch = ""
    if not fp32_in and nq <= 16:
        pk = evdr.PackedPages.from_padded(P, pm, varlen=varlen, page_offset=123)
        full = pk.score(Q, qm)
        k = int(rng.choice([1, 10, 32]))
        ts, ti = pk.search(Q, qm, k)
        ws, wi = evdr.top_k(full, k, id_offset=123)
        same = torch.equal(ti, wi) and torch.equal(ts, ws)
        ok = ok and same
        srch = f" search(k={k}) {'ok' if same else 'MISMATCH'}"
    n_cases += 1
    print(f"[{n_cases:3d}] nq={nq:4d} np={npg:6d} lp={lp:5d} varlen={int(varlen)} fp32={int(fp32_in)} seed={seed:10d} "
          f"err {err:.2e} / {err2:.2e}{srch} {'ok' if ok else 'FAIL'}", flush=True)
    if not ok:
        sys.exit(1)
    del P, Q, pm, qm, got
print(f"stress: {n_cases} cases passed")
