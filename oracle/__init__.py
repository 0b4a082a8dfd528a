"""CPU oracle for the window-attention SR hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product path (``tpu_superresolution_b200``) never
imports this package and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so
the pins are outputs of the *unmodified reference modules* imported in the
build container through ``oracle/reference_loader.py`` and committed under
``tests/golden/`` by ``oracle/make_golden.py``.  ``tests/test_oracle_golden.py``
checks this restatement against those fixtures on every run.
"""
