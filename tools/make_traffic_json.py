#!/usr/bin/env python
"""profiles/traffic.json (what bench.py reports as roofline.traffic) from the committed ncu summaries:
dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, keyed '<queries>x<pages>'.

    python tools/make_traffic_json.py          # reads profiles/r2_maxsim_*.txt written by tools/ncu_summary.py
"""
import json
import re
from pathlib import Path

PROF = Path(__file__).resolve().parent.parent / "profiles"
SOURCES = {"1024x100000": "r2_maxsim_pair8_1024x100k.txt", "64x100000": "r2_maxsim_pair8_64x100k.txt",
           "32x100000": "r2_maxsim_pair8_32x100k.txt", "1x100000": "r2_maxsim_scan_1x100k.txt"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def metric(text, name):
    m = re.search(rf"^{re.escape(name)}\s+([0-9.]+)\s+(\w+)", text, re.M)
    return float(m.group(1)) * UNIT[m.group(2)]


out = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (maxsim_tc_kernel pair "
                "flavour; maxsim_scan_kernel for one query), ncu --set full of the final binary, keyed queries x pages_per_gpu"}
for key, fn in SOURCES.items():
    t = (PROF / fn).read_text()
    kern = re.search(r"^== (.*)$", t, re.M).group(1)
    out[key] = {"bytes": int(metric(t, "dram__bytes_read.sum") + metric(t, "dram__bytes_write.sum")),
                "kernel": kern.split("(")[0], "source": f"profiles/{fn}"}
(PROF / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
