"""In-tree build of libevdr.so with plain nvcc for sm_100a (no torch extension machinery:
the library has a C ABI and is loaded with ctypes).

    python efficient-visual-document-retrieval_b200/csrc/build.py [--force] [--verbose]

Objects go to csrc/build/, the shared library next to the package's __init__.py so that it
travels with the repo snapshot to the GPU box.  cudart is linked statically and neither
libcuda nor libnccl is a link-time dependency (driver entry points and NCCL are resolved at
run time), so the library also loads on a machine without a GPU or driver.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
OUT = PKG / "libevdr.so"
BUILD = CSRC / "build"
SOURCES = ["evdr_core.cu", "maxsim_simt.cu", "pack.cu", "topk.cu", "sharded.cu", "losses.cu", "train.cu", "maxsim_tc.cu", "maxsim_scan.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "topk_dev.cuh", "../../include/evdr.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libevdr.so cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS + ["build.py"]:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = BUILD / "stamp"
    digest = _digest()
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == digest:
        return OUT
    BUILD.mkdir(exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str):
        obj = BUILD / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {src}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(str(obj))
    link = [nvcc, "-shared", "-o", str(OUT), *objs, "-cudart", "static",
            "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libevdr.so failed")
    stamp.write_text(digest)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
