"""pacmensl_b200 -- B200-native FSP time-stepping core (drop-in for PACMENSL's matrix action and ODE
integration).  The compute path is libpacmensl_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/fsp_b200.h); this package holds the Python bindings used by tests and bench.py."""
from . import _capi  # noqa: F401

__all__ = ["_capi"]
