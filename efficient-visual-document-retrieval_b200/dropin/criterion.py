"""``criterion`` drop-in: the seven losses of the reference's criterion.py by the same names, bound to the
fused kernels of libevdr.so (``from criterion import listwise_distillation_loss, score_preserving_loss``
in mainv2_iter_liscore.py:23 etc. resolves here when ``dropin/`` precedes the reference on PYTHONPATH)."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parents[2])
if _root not in sys.path:
    sys.path.insert(0, _root)
_m = importlib.import_module("efficient-visual-document-retrieval_b200.losses")

infonce_supervised_loss = _m.infonce_supervised_loss
infonce_distillation_loss = _m.infonce_distillation_loss
score_preserving_loss = _m.score_preserving_loss
pairwise_distillation_loss = _m.pairwise_distillation_loss
listwise_distillation_loss = _m.listwise_distillation_loss
lambda_loss = _m.lambda_loss
ranknce_loss = _m.ranknce_loss

__all__ = ["infonce_supervised_loss", "infonce_distillation_loss", "score_preserving_loss",
           "pairwise_distillation_loss", "listwise_distillation_loss", "lambda_loss", "ranknce_loss"]
