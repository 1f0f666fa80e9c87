"""Run the tcgen05 MaxSim kernel for a few seconds per configuration while sampling nvidia-smi
(SM clock, power, throttle reasons): separates "pipeline-bound" from "power-capped".
    python tools/tc_power.py [pages] [nq] [seconds]"""
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

pages = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
dev = torch.device("cuda", 0)
corpus = evdr.synth.make_uniform_corpus(pages, 1040, device=dev, seed=1)
pk = evdr.PackedPages.from_uniform_tokens(corpus)
Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, device=dev), dim=-1)
qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
Qb, qm32 = evdr.pack_queries(Q, qm)
out = torch.empty((nq, pages), dtype=torch.float32, device=dev)


def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                          "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append(line.strip())
        if stop.is_set():
            break
    p.terminate()


VARIANTS = [("1cta-ppi16", {"EVDR_TC_MODE": "1cta", "EVDR_TC_DEBUG": "0", "EVDR_TC_PPI": "16"}),
            ("1cta-ppi64", {"EVDR_TC_MODE": "1cta", "EVDR_TC_DEBUG": "0", "EVDR_TC_PPI": "64"}),
            ("pair-ppi16", {"EVDR_TC_MODE": "pair", "EVDR_TC_DEBUG": "0", "EVDR_TC_PPI": "16"}),
            ("pair-ppi64", {"EVDR_TC_MODE": "pair", "EVDR_TC_DEBUG": "0", "EVDR_TC_PPI": "64"}),
            ("1cta-noepi16", {"EVDR_TC_MODE": "1cta", "EVDR_TC_DEBUG": "1", "EVDR_TC_PPI": "16"}),
            ("pair-noepi16", {"EVDR_TC_MODE": "pair", "EVDR_TC_DEBUG": "1", "EVDR_TC_PPI": "16"})]
if os.environ.get("TC_POWER_VARIANTS"):
    import json
    VARIANTS = json.loads(os.environ["TC_POWER_VARIANTS"])
for label, env in VARIANTS * int(os.environ.get("TC_POWER_REPEAT", "2")):
    os.environ.update(env)
    for _ in range(2):
        pk.score_packed(Qb, qm32, out=out)
    torch.cuda.synchronize()
    rows, stop = [], threading.Event()
    th = threading.Thread(target=sample, args=(stop, rows), daemon=True)
    th.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        pk.score_packed(Qb, qm32, out=out)
        n += 1
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    stop.set()
    time.sleep(0.25)
    clk, pw, tmp, cap = [], [], [], 0
    for r in rows[3:]:
        f = [x.strip() for x in r.split(",")]
        try:
            clk.append(float(f[0])); pw.append(float(f[1])); tmp.append(float(f[2]))
            cap += f[3].lower().startswith("active")
        except Exception:
            pass
    tf = nq * pages * 2 * 128 * 32 * 1030 / (ms * 1e-3) / 1e12
    print(f"{label:12s} {ms:8.2f} ms  {tf:7.1f} TFLOP/s  clk med {statistics.median(clk) if clk else 0:6.0f} MHz "
          f"min {min(clk) if clk else 0:5.0f}  power med {statistics.median(pw) if pw else 0:6.0f} W  temp {max(tmp) if tmp else 0:.0f}C "
          f"power_cap {cap}/{len(clk)}  util/clk {tf*1e12/(148*8192*statistics.median(clk)*1e6) if clk else 0:.3f}", flush=True)
    time.sleep(1.0)
