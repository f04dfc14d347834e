"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the Gleipnir built-in Nested Sampling hot path, used as the
parity checker by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under ``gleipnir_b200/`` or
``gleipnir/`` may import this package.
"""
