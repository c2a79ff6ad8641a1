"""CPU oracle for the TS-SEP inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker.  The product
(``tssep_b200``) never imports this package and has no CPU fallback.
"""
