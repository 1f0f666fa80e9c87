"""CPU oracle for the late-interaction (MaxSim) scoring path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product package never imports this directory and
fails loudly when its CUDA extension is missing.

Parity status
-------------
* Scoring half (``maxsim_oracle``): PINNED.  The restatement is checked against
  outputs of the unmodified reference function
  ``evaluator/retrieval.py:166-213`` executed in the build container
  (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``), including the
  known-answer micro-cases of SURVEY.md section 4.
* Metric half (``trec_oracle``): PARITY UNPINNED.  The reference delegates to
  ``mteb`` -> ``pytrec_eval`` (``evaluator/retrieval.py:239-246``), neither of
  which is vendored, pinned or installed.  The restatement follows trec_eval's
  published definitions (ndcg_cut, map_cut, recall, P) and mteb 1.x's
  mean + round(5); it is cross-checked against scikit-learn's ``ndcg_score``
  on tie-free cases only.
"""
