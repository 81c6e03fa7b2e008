"""CPU oracle for the Fit-Hi-C significance pass.  TEST INFRASTRUCTURE ONLY.

Nothing in ``blueberry_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` use it, and only as the checker or the timed CPU baseline.

Parity status
-------------
The reference (jmschrei/blueberry) ships no tests, golden vectors or fixtures
for this path, so the oracle is pinned the only way available: by EXECUTING the
reference's own code in the build container (``oracle/ref_loader.py`` runs
``/root/reference/blueberry/fithic.py`` after the closed list of Python-2 -> 3
mechanical edits, and compiles ``blueberry.pyx`` verbatim into
``oracle/_ref/``) and freezing the outputs under ``tests/golden/`` next to the
script that made them (``oracle/make_golden.py``).  The third-party arithmetic
the reference delegates to (scipy.special.bdtrc, scipy UnivariateSpline,
sklearn IsotonicRegression) is pinned only by the versions installed in this
image (scipy 1.18.1, scikit-learn 1.9.0, numpy 2.3.5) - the reference itself
pins nothing.
"""
