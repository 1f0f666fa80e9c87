"""Is a kernel pipeline-bound or power-bound?  The same launches are timed (CUDA events, per launch) twice:
back to back for a few seconds (the power cap engages and the SM clock settles where ~1000 W are met), and
duty-cycled -- every launch followed by an idle gap of `gap` launch-times, so that the average power stays
below the cap and the clock stays high.  A pipeline-bound kernel runs equally fast (per launch) in both; a
power-bound one is faster when duty-cycled, by about the ratio of the SM clocks.

    python tools/power_trace.py [pages] [seconds] [nq ...]        # default 100000 3.0 1 16 32 64 1024
"""
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

pages = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
nqs = [int(x) for x in sys.argv[3:]] or [1, 16, 32, 64, 1024]
dev = torch.device("cuda", 0)
corpus = evdr.synth.make_uniform_corpus(pages, 1040, device=dev, seed=1)
pk = evdr.PackedPages.from_uniform_tokens(corpus)
FLOPS_PER_PAIR, BYTES_PER_PAGE = 2 * 128 * 32 * 1030, 1030 * 128 * 2


class Smi:
    def __init__(self):
        self.rows, self.stop = [], threading.Event()
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        clk, pw, cap = [], [], 0
        for t, r in self.rows:
            if t0 <= t <= t1:
                f = [x.strip() for x in r.split(",")]
                try:
                    clk.append(float(f[0])); pw.append(float(f[1])); cap += f[2].lower().startswith("active")
                except Exception:
                    pass
        return clk, pw, cap


smi = Smi()
time.sleep(0.5)
for nq in nqs:
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=dev), dim=-1)
    qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
    Qb, qm32 = evdr.pack_queries(Q, qm)
    out = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    for _ in range(3):
        pk.score_packed(Qb, qm32, out=out)
    torch.cuda.synchronize()
    time.sleep(1.0)
    res = {}
    for mode, gap in (("back-to-back", 0.0), ("duty-cycled", 4.0)):
        evs, t0 = [], time.perf_counter()
        while time.perf_counter() - t0 < secs:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pk.score_packed(Qb, qm32, out=out)
            e1.record()
            evs.append((e0, e1))
            if gap:
                torch.cuda.synchronize()
                time.sleep(gap * e0.elapsed_time(e1) * 1e-3)
            elif len(evs) % 64 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = [a.elapsed_time(b) for a, b in evs[len(evs) // 3:]]        # the settled part
        clk, pw, cap = smi.window(t0 + (t1 - t0) / 3, t1)
        med = statistics.median(ms)
        res[mode] = med
        tf = nq * pages * FLOPS_PER_PAIR / (med * 1e-3) / 1e12
        gbs = pages * BYTES_PER_PAGE / (med * 1e-3) / 1e9
        print(f"nq={nq:5d} {mode:13s} {len(evs):5d} launches  median {med:9.3f} ms  {tf:7.1f} TFLOP/s  {gbs:7.1f} GB/s of pages  "
              f"SM clock med {statistics.median(clk) if clk else 0:5.0f} MHz (min {min(clk) if clk else 0:4.0f})  "
              f"power med {statistics.median(pw) if pw else 0:5.0f} W  power_cap {cap}/{len(clk)} samples", flush=True)
        time.sleep(1.5)
    print(f"nq={nq:5d} duty-cycled / back-to-back speed: {res['back-to-back'] / res['duty-cycled']:.3f}x", flush=True)
smi.p.terminate()
