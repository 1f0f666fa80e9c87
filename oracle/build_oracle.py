"""Compile oracle/maxsim_oracle.c -> oracle/_build/liboracle.so with gcc (TEST INFRASTRUCTURE).
The reference itself is pure Python: there is nothing to compile into oracle/_ref."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_build" / "liboracle.so"


def build(force: bool = False) -> Path:
    src = HERE / "maxsim_oracle.c"
    if OUT.exists() and not force and OUT.stat().st_mtime >= src.stat().st_mtime:
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-o", str(OUT), str(src)], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
