"""``evaluator.retrieval`` drop-in: re-exports the B200 mirror of evaluator/retrieval.py."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parents[3])
if _root not in sys.path:
    sys.path.insert(0, _root)
_m = importlib.import_module("efficient-visual-document-retrieval_b200.retrieval")

score_multi_vector_masked = _m.score_multi_vector_masked
BaseVisualRetrieverProcessor = _m.BaseVisualRetrieverProcessor
CustomRetrievalEvaluator = _m.CustomRetrievalEvaluator
get_torch_device = _m.get_torch_device
left_padding = _m.left_padding

__all__ = ["score_multi_vector_masked", "BaseVisualRetrieverProcessor", "CustomRetrievalEvaluator",
           "get_torch_device", "left_padding"]
