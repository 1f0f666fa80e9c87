#!/usr/bin/env python
"""Stage the UNMODIFIED reference files of the hot path into the git-ignored ``baseline/_ref/``.

    python tools/stage_reference.py            # called by __graft_entry__.build() in the build container

``/root/reference`` only exists in the build container; ``baseline/_ref/`` is git-ignored (reference
sources never enter this repository's history) but it is NOT gpurun-ignored, so the staged files travel
with the snapshot to the GPU box.  There they serve three purposes, all outside the product path:

* ``bench.py --impl reference`` and ``cpu_baseline``: the reference's own ``score_multi_vector_masked`` timed
  on the host cores (``kind: "reference"``);
* ``bench.py`` ``gpu_reference``: the same function torch-eager on the B200;
* ``tests/``: the comparator of the config-shaped GPU tests and the script acceptance test
  (``train_one_step`` / ``eval_retrieval`` of ``mainv2_iter_liscore.py`` run unmodified against the drop-in).

Files are copied byte for byte; ``MANIFEST.json`` records their sha256 so that ``baseline/ref.py`` can refuse
a staged tree that was edited.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DEST = ROOT / "baseline" / "_ref"
SRC = Path(os.environ.get("EVDR_REFERENCE_SRC", "/root/reference"))

# the scorer + losses, what they import, and the scripts whose step / eval functions the acceptance tests run
FILES = [
    "evaluator/retrieval.py",
    "criterion.py",
    "utils/preprocess_data.py",
    "utils/utils.py",
    "utils/mapping.py",
    "Qdatasets/query_tensor_dataset.py",
    "mainv2_iter_liscore.py",
    "mainv2_iter_listwise.py",
    "mainv2_iter_ranknet.py",
    "mainv3_iter_liscore_QA_hardtoken.py",
    "mainv3_iter_liscore_mixup.py",
]


def sha256(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()


def stage(verbose: bool = True) -> bool:
    """Copy FILES from the reference checkout; returns False when there is no checkout (GPU box)."""
    if not SRC.is_dir():
        if verbose:
            state = "present" if (DEST / "MANIFEST.json").exists() else "ABSENT"
            print(f"[stage_reference] {SRC} not found; staged copy is {state}")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = SRC / rel, DEST / rel
        if not src.is_file():
            raise FileNotFoundError(f"reference file missing: {src}")
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
        assert manifest[rel] == sha256(src)
    (DEST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "files": manifest}, indent=1))
    if verbose:
        print(f"[stage_reference] staged {len(manifest)} unmodified files into {DEST}")
    return True


if __name__ == "__main__":
    ok = stage()
    sys.exit(0 if ok or (DEST / "MANIFEST.json").exists() else 1)
