"""CPU oracle for the main_network_best training step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the product package
(``depth-enhancement-and-super-resolution_b200/``).  The only legal importers are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

The oracle restates, in plain CPU PyTorch / numpy, the algorithm of the reference's hot path
(``/root/reference/models/main_model.py`` and the files it calls).  Every function cites the
reference file:line it follows.

Parity pin: the reference ships no golden vectors or tests, so the oracle is
pinned against OUTPUTS OF THE LIVE REFERENCE, imported in the build container from
``/root/reference`` by ``tests/golden/make_golden.py`` (committed) which wrote the fixtures under
``tests/golden/*.npz``.  ``tests/test_oracle_vs_golden.py`` checks the oracle against those
fixtures on every run; ``tests/test_oracle_vs_reference.py`` re-checks against the live reference
whenever ``/root/reference`` is present (it is not on the GPU box).
"""
