"""Importable alias of the package directory ``efficient-visual-document-retrieval_b200`` (whose
name is not a Python identifier).  ``import evdr_b200`` and ``import evdr_b200.retrieval`` resolve
to the very same module objects."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent.parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = importlib.import_module("efficient-visual-document-retrieval_b200")
for _sub in _real.SUBMODULES:
    sys.modules[f"{__name__}.{_sub}"] = getattr(_real, _sub)
sys.modules[__name__] = _real
