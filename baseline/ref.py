"""Loader for the UNMODIFIED reference staged under ``baseline/_ref/`` (see tools/stage_reference.py).

Baseline / test infrastructure, never imported by the product package.  Everything here executes the
reference's own source files byte for byte (sha256-checked against ``MANIFEST.json``); the only thing added is
an empty ``mteb`` module tree in ``sys.modules`` because ``evaluator/retrieval.py:218`` imports mteb's
RetrievalEvaluator at module level and mteb is not installable offline.  The metric
half of the reference (``compute_mteb_metrics``) therefore cannot run; callers pass their own evaluator
object to the scripts' ``eval_retrieval(evaluator=...)``.
"""
from __future__ import annotations

import contextlib
import hashlib
import importlib
import importlib.util
import json
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
DROPIN = ROOT / "efficient-visual-document-retrieval_b200" / "dropin"

# top-level names the reference's scripts import; they are swapped in and out of sys.modules around a load so
# that a stock load and a drop-in load can live in one process
_REF_TOPLEVEL = ("evaluator", "criterion", "utils", "Qdatasets")
_verified = False


def available() -> bool:
    return (REF / "MANIFEST.json").is_file()


def why_unavailable() -> str:
    return (f"{REF} is not staged: run `python tools/stage_reference.py` (or __graft_entry__.build()) in a "
            f"container that has /root/reference")


def verify() -> dict:
    """sha256 of every staged file against the manifest written at staging time."""
    global _verified
    man = json.loads((REF / "MANIFEST.json").read_text())
    if not _verified:
        for rel, digest in man["files"].items():
            got = hashlib.sha256((REF / rel).read_bytes()).hexdigest()
            if got != digest:
                raise RuntimeError(f"staged reference file {rel} differs from the manifest: the baseline must run "
                                   f"the UNMODIFIED reference")
        _verified = True
    return man


def _stub_mteb() -> None:
    for name in ("mteb", "mteb.evaluation", "mteb.evaluation.evaluators",
                 "mteb.evaluation.evaluators.RetrievalEvaluator"):
        sys.modules.setdefault(name, types.ModuleType(name))
    leaf = sys.modules["mteb.evaluation.evaluators.RetrievalEvaluator"]
    if not hasattr(leaf, "RetrievalEvaluator"):
        leaf.RetrievalEvaluator = type("RetrievalEvaluator", (), {})


@contextlib.contextmanager
def _import_scope(dropin: bool):
    """sys.path = [dropin/]? + [baseline/_ref] + rest, with the reference's top-level module names cleared;
    both are restored on exit (modules imported inside the scope stay referenced by whoever imported them)."""
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_TOPLEVEL}
    for k in saved:
        del sys.modules[k]
    old_path = list(sys.path)
    sys.path[:0] = ([str(DROPIN)] if dropin else []) + [str(REF)]
    importlib.invalidate_caches()
    _stub_mteb()
    try:
        yield
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_TOPLEVEL]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = old_path
        importlib.invalidate_caches()


def _load_file(rel: str, alias: str):
    spec = importlib.util.spec_from_file_location(alias, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


_cache: dict = {}


def scorer_module():
    """The reference's ``evaluator/retrieval.py`` (``score_multi_vector_masked``, ``BaseVisualRetrieverProcessor``)."""
    if "scorer" not in _cache:
        if not available():
            raise FileNotFoundError(why_unavailable())
        verify()
        _stub_mteb()
        _cache["scorer"] = _load_file("evaluator/retrieval.py", "_evdr_reference_evaluator_retrieval")
    return _cache["scorer"]


def criterion_module():
    """The reference's ``criterion.py``."""
    if "criterion" not in _cache:
        if not available():
            raise FileNotFoundError(why_unavailable())
        verify()
        _cache["criterion"] = _load_file("criterion.py", "_evdr_reference_criterion")
    return _cache["criterion"]


def preprocess_module():
    """The reference's ``utils/preprocess_data.py`` (l2_normalize, preprocess_docs, preprocess_queries)."""
    if "preprocess" not in _cache:
        if not available():
            raise FileNotFoundError(why_unavailable())
        verify()
        _cache["preprocess"] = _load_file("utils/preprocess_data.py", "_evdr_reference_preprocess_data")
    return _cache["preprocess"]


def load_script(script: str, dropin: bool):
    """Execute a staged ``main*.py`` as a module (its ``main()`` is guarded by ``__name__``) and return it.

    dropin=False: its ``from evaluator.retrieval import ...`` / ``from criterion import ...`` bind to the
    reference's own files.  dropin=True: ``dropin/`` precedes the reference on sys.path exactly as
    INTEGRATION.md prescribes (PYTHONPATH=dropin), so the same unmodified script binds to this repository's
    scorer, evaluator and losses.  The two loads are independent module objects."""
    if not available():
        raise FileNotFoundError(why_unavailable())
    verify()
    alias = f"_evdr_reference_script_{Path(script).stem}_{'dropin' if dropin else 'stock'}"
    with _import_scope(dropin):
        mod = _load_file(script, alias)
        bound = sys.modules["evaluator.retrieval"].__file__
    want = DROPIN if dropin else REF
    if not Path(bound).resolve().is_relative_to(want.resolve()):
        raise RuntimeError(f"{script}: evaluator.retrieval bound to {bound}, expected something under {want}")
    return mod
