"""Diagnostic for the tcgen05 MaxSim kernel: each configuration runs in its own process (a device
trap poisons the CUDA context) with a timeout, and prints the error against the exact GPU kernel.

    python tools/tc_diag.py            # all configurations
    python tools/tc_diag.py one NQ NP LP [PPI]
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CONFIGS = [  # nq, np, lp, ppi
    (4, 1, 256, 0), (4, 1, 16, 0), (4, 2, 1030, 0), (1, 16, 1030, 0), (4, 64, 1030, 0), (8, 64, 1030, 0),
    (12, 64, 1030, 0), (13, 200, 1030, 0), (40, 500, 1030, 0), (20, 4000, 64, 0), (500, 500, 1030, 0),
    (64, 20000, 1030, 0),
]


def one(nq, npg, lp, ppi):
    import time
    import torch
    import evdr_b200 as evdr
    dev = torch.device("cuda", 0)
    P, pm = evdr.synth.make_pages(npg, lp, seed=1, device=dev, n_holes=0, empty_frac=0.0)
    Q, qm, rel = evdr.synth.make_queries(P, pm, nq, 32, seed=2, sigma=2.0)
    pk = evdr.PackedPages.from_padded(P, pm)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = pk.score_packed(Qb, qm32, pages_per_item=ppi)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if nq * npg * lp <= 40 * 500 * 1030 * 4:
        ref = evdr.score_multi_vector_masked(Q, P, qm, pm, precision="fp32")
    else:
        sub = slice(0, 200)
        ref = torch.full_like(got, float("nan"))
        ref[:8, sub] = evdr.score_multi_vector_masked(Q[:8], P[sub], qm[:8], pm[sub], precision="fp32")
    mask = ~torch.isnan(ref)
    err = ((got - ref).abs() / ref.abs().clamp_min(1.0))[mask]
    print(f"nq={nq} np={npg} lp={lp} ppi={ppi}: max rel err {err.max().item():.3e}  first-call {dt*1e3:.2f} ms"
          f"  got[0,:4]={got[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}", flush=True)
    if err.max().item() > 1e-3:
        bad = torch.nonzero(((got - ref).abs() / ref.abs().clamp_min(1.0)) > 1e-3)
        print("   mismatches:", bad.shape[0], "first:", bad[:8].tolist(), flush=True)
        sys.exit(3)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        a = [int(x) for x in sys.argv[2:]]
        one(a[0], a[1], a[2], a[3] if len(a) > 3 else 0)
    else:
        fails = 0
        for cfg in CONFIGS:
            try:
                r = subprocess.run([sys.executable, __file__, "one", *map(str, cfg)], timeout=180,
                                   capture_output=True, text=True)
                print(r.stdout.strip())
                if r.returncode != 0:
                    fails += 1
                    print(f"   rc={r.returncode} stderr tail: {r.stderr.strip()[-600:]}")
            except subprocess.TimeoutExpired:
                fails += 1
                print(f"   TIMEOUT {cfg}")
        print("diag failures:", fails)
