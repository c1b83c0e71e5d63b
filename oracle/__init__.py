"""CPU oracle for the SGPR prediction hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker / the CPU arm -- never on the
product path (``autoforce_b200`` fails loudly when its CUDA library is missing).
"""
