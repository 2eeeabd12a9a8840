"""pulsarbat_b200 -- B200 (sm_100a) kernels behind pulsarbat's FFT baseband hot path.

The compute lives in libpbk.so (include/pbk.h); this package is the thin Python mirror of the
reference's interface for that path.  There is no CPU fallback.
"""

from . import _lib  # noqa: F401
from ._lib import PbkError, PbkUnsupported  # noqa: F401

__version__ = "0.1.0"
