"""CPU oracle for the fused joint + RNN-T loss path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``rnntransducer_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do.  See ``oracle/warp_cpu.c`` for provenance and the parity pin.
"""
