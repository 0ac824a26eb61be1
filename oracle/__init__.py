"""CPU oracle for the qBOLD-VI hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, op for op, the arithmetic of the reference's
``signals.py`` (forward ASE qBOLD signal model) and the sampling / likelihood /
KL slice of ``model.py`` in NumPy, in two flavours:

* ``np.float32`` -- every elementary operation rounded to float32, Bessel
  functions through the Cephes single-precision kernels that TensorFlow's
  ``tf.math.special.bessel_j0/j1`` resolve to (Eigen ``generic_j0/j1<float>``);
* ``np.float64`` -- the same formulas in double precision, with the one FP32
  artefact that is part of the reference's results (quadrature node 0 is dead in
  the value, live in the derivative) reproduced explicitly.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import anything from here, and only as the checker.  The product path
(``qbold_vi_b200``) never imports it and has no CPU fallback.

PARITY PINNING STATUS: the reference ships no tests, golden vectors or fixtures
 and TensorFlow cannot be installed offline, so the oracle
is *not* pinned against real TensorFlow output ("parity unpinned" w.r.t. TF).
It IS pinned against the reference's own, unmodified source files executed over
a torch-backed TensorFlow API shim (``oracle/tf_shim`` + ``oracle/make_golden.py``
-> ``tests/golden/ref_shim_*.npz``) and against the known-answer vectors of
SURVEY.md Appendix B (``tests/golden/kat_appendix_b.npz``).
"""
