"""CPU oracle for the regularizepsf correction hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``regularizepsf_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may use it, and only as the checker or the timed CPU baseline.
"""
