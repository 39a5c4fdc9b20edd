"""destr-b200: B200-native (sm_100a) implementation of the DESTR transformer-half hot path.

Importing the package loads libdestr_b200.so (built by __graft_entry__.build()); it raises
ImportError when the library is missing -- there is no CPU or pure-torch fallback.
"""
from . import _lib  # noqa: F401  (fails loudly without the CUDA library)
from . import ops  # noqa: F401

__all__ = ["ops"]
