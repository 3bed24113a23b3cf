"""CPU oracle for the post-detection box pipeline (soft-NMS ensemble + SORT).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or as the timed CPU baseline.  The product
(``waymo_2d_tracking_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Parity status
-------------
* Ensemble side (``ensemble_port``): PINNED.  The reference's own files
  (``detnet/ensemble.py``, ``detnet/nn/tta.py``, ``detnet/utils/box_utils.py``)
  run in the dev container through ``ref_shim``; ``tests/golden/make_golden.py``
  executed them to produce the committed fixtures and the port reproduces those
  bit for bit.
* SORT side (``sort_port``): the reference's own ``tracking/utils.py``,
  ``tracking/sort/{sort,tracker_sort}.py`` also run through ``ref_shim``, but two
  third-party pieces they import are absent from ``/root/reference`` and from
  this image: scikit-learn 0.22.2 ``sklearn.utils.linear_assignment_``
  (``environment.yml:18``) and ``filterpy.kalman.KalmanFilter`` (unpinned,
  ``environment.yml:35``).  They are restated from their published algorithms
  in ``munkres.py`` / ``kalman.py``.  The reference holds no tests or golden
  vectors for this path, so at those two boundaries the parity
  is **unpinned**: the restatement is the definition, cross-checked for
  optimal cost against SciPy's ``linear_sum_assignment`` and for the Kalman
  covariance identities.  Everything around them is pinned by executing the
  reference's own code.

NumPy semantics: two promotion regimes, selected per call (``promotion="legacy"``
— NumPy 1.x value-based casting, the reference's pinned environment and the
product's default — or ``"nep50"``, NumPy 2).  The ports use explicit casts so they
do not depend on the installed NumPy's promotion rules.  The ``nep50`` goldens are
the reference's files executed as they are under this image's NumPy 2; for
``legacy`` the two promotion-sensitive spots of that execution are emulated
(``ref_shim._legacy_convert_bbox_to_z`` and float64 thresholds), because NumPy 1.x
cannot be installed here — see DESIGN.md §3.
"""
