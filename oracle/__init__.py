"""CPU oracle for the ADER hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``ader_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

What it restates (reference = doublemul/ADER, paths relative to /root/reference):

* ``sasrec.py``   -- modules.py:23-271 (LayerNorm, embedding, attention, FFN),
                     ADER.py:25-150 (assembly, rep, logits, CE, KD / ER loss, ranking),
                     EWC.py:115-164 (penalty, Fisher), TF1 Adam (tf.train.AdamOptimizer).
* ``protocol.py`` -- util.py:17-522 (DataLoader, Sampler, Evaluator.results,
                     ExemplarGenerator quota / herding / loss / random selection),
                     main.py:54-65,181-201 (exemplar flattening, lambda schedule).

Parity pinning.  The reference ships no tests, golden vectors or logs,
and its numeric core is TensorFlow 2.1, which is not installable here.  So:

* ``protocol.py`` IS pinned: ``tests/golden/make_golden.py`` runs the reference's own
  ``util.py`` (unmodified, behind a stub ``tensorflow`` module) in the build container and
  commits its outputs (sampler batches, valid split, herding picks, multinomial quotas,
  metrics) as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
  this restatement against them (and against the live reference when it is mounted).
* ``sasrec.py`` is **parity unpinned** against TensorFlow itself: it is a line-by-line
  restatement of the TF graph in torch-CPU (fp32, with an fp64 twin), self-checked by
  finite differences and by dense-vs-packed equivalence, but no TF-produced vector exists
  to pin it.  DESIGN.md says the same.
"""
