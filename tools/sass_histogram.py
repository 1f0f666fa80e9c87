#!/usr/bin/env python
"""Static Blackwell evidence: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use.

    python tools/sass_histogram.py [libevdr.so] > profiles/sass_histogram.txt

Runs `cuobjdump -sass` on the built library (no GPU needed) and counts, per kernel,
UTCHMMA (tcgen05.mma, incl. the .2CTA form), LDTM / STTM (tcgen05.ld / st), UTMALDG (cp.async.bulk.tensor),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), FMNMX3 (3-input max of the epilogue), REDUX, plus the
instruction count.  Kernels without any of the tensor / TMA mnemonics are summarised in one line each.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "efficient-visual-document-retrieval_b200" / "libevdr.so"
PATTERNS = collections.OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM\b"),
    ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"), ("UTCBAR", r"\bUTCBAR\b"), ("SYNCS", r"\bSYNCS\b"),
    ("FMNMX3", r"\bFMNMX3\b"), ("REDUX", r"\b(CREDUX|REDUX)\b"), ("HFMA2/FFMA", r"\b(FFMA|HFMA2)\b"),
])


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def shorten(name):
    """Drop the parameter list, keep the template arguments (`(int)2` -> `2`, `(bool)1` -> `1`)."""
    name = re.sub(r">\(.*$", ">", name) if ">(" in name else re.sub(r"\(.*$", "", name)
    return re.sub(r"\((?:int|bool)\)", "", name)


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("/*")[1] if line.strip().startswith("/*") else line
        if not re.search(r"/\*[0-9a-f]{4}\*/", line):
            continue
        cur["instructions"] += 1
        for key, pat in PATTERNS.items():
            if re.search(pat, body):
                cur[key] += 1
    names = demangle(list(kernels))
    print(f"# cuobjdump -sass {LIB.name}: architectures {', '.join(arch)}; {len(kernels)} kernels")
    print("# per kernel: " + " ".join(PATTERNS) + " | instructions")
    tot = collections.Counter()
    tensor, other = [], []
    for k, c in kernels.items():
        tot.update(c)
        (tensor if c["UTCHMMA"] + c["UTCHMMA.2CTA"] + c["UTMALDG"] + c["LDTM"] else other).append((k, c))
    print(f"\n## kernels using tcgen05 / TMEM / TMA ({len(tensor)})")
    for k, c in tensor:
        short = shorten(names.get(k, k))
        print(f"{short}\n    " + "  ".join(f"{key}={c[key]}" for key in PATTERNS) + f"  | {c['instructions']}")
    print(f"\n## other kernels ({len(other)})")
    for k, c in other:
        short = shorten(names.get(k, k))
        print(f"{short}: {c['instructions']} instructions, FMNMX3={c['FMNMX3']} REDUX={c['REDUX']} FFMA/HFMA2={c['HFMA2/FFMA']}")
    print("\n## totals\n" + "  ".join(f"{key}={tot[key]}" for key in PATTERNS) + f"  | {tot['instructions']} instructions")


if __name__ == "__main__":
    main()
