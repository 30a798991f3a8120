"""CPU oracle for the pairwise-ranking hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference's hot path
(BinFuPKU/CollaborativeFilteringUsingTensorflow: src/models/pl/models/{bprmf,cml,gbprmf}.py,
src/models/basic/models/{wrmf,mf,svd}.py, src/samplers/sampler_*.py, src/metrics/{ranking,rating}.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed CPU
baseline.  The product package ``collaborativefilteringusingtensorflow_b200`` never
imports it and has no CPU fallback.

Pinning status
--------------
* ``oracle.ranking``  -- PINNED: checked against the reference's own ``ranking.py`` run
  live in the build container and against the golden vectors captured from its
  ``__main__`` toy inputs (tests/golden/ranking_golden.json, made by oracle/gen_golden.py).
* ``oracle.scoring.topn_masked`` + ``oracle.ranking`` end to end -- PINNED against the OUTPUT of the
  reference's own PopRank (basic/models/pop.py, numpy only) run live on ml-100k fold 1: its recommended
  lists and metric values (tests/golden/pop_golden.json).
* ``oracle.samplers`` -- PINNED on invariants/dtypes/shapes against the reference samplers
  run live (tests/golden/sampler_golden.json); streams are unseeded in the reference so no
  stream-level golden vectors exist.
* ``oracle.rating``   -- metrics PINNED against the reference's own ``metrics/rating.py`` run live
  (tests/golden/rating_golden.json); the MF step / prediction restatement is unpinned like
  ``oracle.steps`` (it IS ``oracle.steps.wrmf_step`` with weight 1); ``oracle.steps.svd_step`` is cross-checked
  against a torch-autograd restatement of svd.py:52-80 (tests/golden/svd_golden.npz).
* ``oracle.steps`` / ``oracle.scoring`` -- **PARITY UNPINNED against the TensorFlow binary**: the
  arithmetic lives in third-party TensorFlow (>=1.13, unpinned, README.md:20-22), which is
  not installable here and for which the reference ships no tests or golden vectors.  Pinned
  instead to (1) the reference's OWN model files, imported unmodified from /root/reference and run
  through their train() on the torch-backed TF-1.x stand-in ``oracle/tf1_shim`` (made by
  oracle/gen_refgraph_golden.py -> tests/golden/{step,tuple,svd}_refgraph_golden.npz: tables,
  accumulators, losses after every step and the metric values train() returned), and (2) an
  independent torch-autograd restatement of the same TF graphs with
  ``torch.optim.Adagrad(initial_accumulator_value=0.1, eps=0)`` (tests/golden/step_golden.npz,
  made by oracle/gen_golden.py).  The stand-in follows TF1's published semantics (SURVEY.md
  Appendix A/B; listed in its header); it is an interpretation of TensorFlow, not TensorFlow.
* ``oracle/tf1_shim`` -- the stand-in itself; imported only by oracle/gen_refgraph_golden.py.
"""
