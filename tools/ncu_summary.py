"""Summarise ncu outputs into small text files for profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof_tc.ncu-rep profiles/r1_maxsim_tc.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source {src}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write(f"{'total ms':>12} {'share':>8} {'launches':>8}  kernel\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t/1e6:12.3f} {100*t/tot:7.2f}% {c:8d}  {n[:140]}\n")


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source {src}\n")
        for V in rows[2:]:
            f.write(f"\n== {V[H.index('Kernel Name')][:160]}\n")
            for k in KEYS:
                if k in H:
                    i = H.index(k)
                    f.write(f"{k:75s} {V[i]:>20s} {U[i]}\n")
            for i, h in enumerate(H):
                if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                    try:
                        if float(V[i]) > 0.05:
                            f.write(f"{h:75s} {V[i]:>20s} {U[i]}\n")
                    except ValueError:
                        pass


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
