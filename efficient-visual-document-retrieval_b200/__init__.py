"""efficient-visual-document-retrieval_b200 -- B200-native late-interaction (MaxSim) scoring.

Drop-in for ONE hot path of kimjy-st/Efficient-Visual-Document-Retrieval: the ColPali-style MaxSim
scorer (evaluator/retrieval.py:166-213), its backward, and the top-k / nDCG@k tail of evaluation.
Host code is PyTorch (device memory, streams, torch.distributed) calling libevdr.so -- hand-written
sm_100a CUDA behind the C ABI of include/evdr.h.  No Triton, no multi-backend dispatch, no CPU
fallback: compute calls raise when the library or a B200 is missing.

The directory name is not a Python identifier; import it with
``importlib.import_module("efficient-visual-document-retrieval_b200")`` or through the alias
package ``evdr_b200`` at the repository root.
"""
from . import _lib
from . import corpus, losses, mining, normalize, retrieval, sharded, synth, topk
from .corpus import HostCorpus, PackedPages, pack_queries
from .losses import (infonce_distillation_loss, infonce_supervised_loss, lambda_loss, liscore_losses,
                     listwise_distillation_loss, pairwise_distillation_loss, rank_gap, ranknce_loss,
                     score_preserving_loss)
from .mining import hard_token_argmax, hard_tokens
from .normalize import l2_normalize, l2_normalize_masked
from .retrieval import (BaseVisualRetrieverProcessor, CustomRetrievalEvaluator, TeacherScoreTable, maxsim_exact,
                        score_multi_vector_masked)
from .sharded import GraphedSearch, ShardedSearch, gather_score_rows, init_nccl_comm, init_peer_exchange, partition_pages
from .topk import retrieval_metrics, topk as top_k, topk_merge, topk_metrics

__version__ = "0.1.0"
SUBMODULES = ("_lib", "corpus", "losses", "mining", "normalize", "retrieval", "sharded", "synth", "topk")
