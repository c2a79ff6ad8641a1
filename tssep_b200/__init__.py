"""tssep_b200: B200-native (sm_100a) implementation of the TS-SEP inference hot path.

Drop-in classes for the reference's ``factory:`` config keys (see
``tssep_b200.configurable.FACTORY_ALIASES``); all arithmetic runs in
``tssep_b200/_lib/libtssep_b200.so`` (build: ``python -m tssep_b200.build``).
"""
__version__ = "0.1.0"
