"""CPU oracle for the SMCP Newton-system hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  Nothing under ``smcp_b200/`` imports it; the
product path fails loudly when the CUDA library is missing.

PARITY UNPINNED.  The arithmetic of this path lives in two third-party dependencies that
are NOT vendored under ``/root/reference`` and are not installable here (no network, not in
the offline wheelhouse):

* ``chompack >= 2.3.4``  (supernodal chordal-matrix kernels: cholesky, completion,
  projected_inverse, llt, hessian, trsm, dot, symbolic) — ``pyproject.toml:25-28``
* ``cvxopt >= 1.3.3``    (BLAS/LAPACK wrappers, sparse gemv, potrf/potrs) — same lines.

Neither is pinned by a lock file, and the reference's own tests assert nothing numerical
about this path (``tests/test_basic.py:19-22`` only prints a status).  The oracle therefore
restates the *published* algorithms (Andersen, Dahl, Vandenberghe, "Implementation of
nonsymmetric interior-point methods for linear optimization over sparse matrix cones",
Math. Prog. Comp. 2010) and is anchored three ways:

1. ``oracle.dense`` — dense NumPy linear algebra (``numpy.linalg``) as ground truth for
   every chordal kernel on small patterns;
2. ``oracle.supernodal`` — the supernodal recursions, checked against (1), used for sizes
   where dense is impractical and as the timed CPU baseline;
3. the reference's own call sites: the IPM drivers in ``smcp_b200.solvers`` follow
   ``src/python/solvers.py`` line by line and run unchanged on this oracle backend and on
   the CUDA backend, so iteration-count parity is purely numerical.
"""
