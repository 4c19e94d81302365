"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ResDepth hot path.

Nothing under ``resdepth_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` use it, and only as the checker / the reported CPU baseline.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the *unmodified* reference classes
(``/root/reference/lib/UNet.py``, ``lib/Trainer.py``, ``lib/evaluation.py``)
run in the build container; the vectors are committed under ``tests/golden/``
together with the generator ``oracle/make_golden.py``.
"""
