"""Shadow of the reference's ``evaluator`` package.  Put ``<repo>/efficient-visual-document-
retrieval_b200/dropin`` in front of the reference checkout on PYTHONPATH and every
``from evaluator.retrieval import score_multi_vector_masked, CustomRetrievalEvaluator``
(mainv1.py:14 ... mainv3_iter_liscore_noisev1.py:20) binds to the B200 implementation."""
