"""CPU oracle for the CNMNet inference hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the timed CPU baseline.
The product path (``cnmnet_b200``) never imports this package and raises if its
CUDA library is missing.

Parity status: the reference (xxlong0/CNMNet) ships no tests, fixtures or golden
vectors ("parity unpinned" by the reference itself).  The oracle is
therefore pinned against outputs of the reference itself: ``oracle/make_golden.py``
imports the unmodified reference modules from /root/reference through the shim in
``oracle/ref_shim.py`` (only possible in the build container), runs them on seeded
synthetic inputs and commits the results under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every restatement here against those vectors.
"""
