"""Counterpart of ``tssep/util``: frame <-> sample activity conversion."""
