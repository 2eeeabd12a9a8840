"""pulsarbat_b200 -- B200 (sm_100a) kernels behind pulsarbat's FFT baseband hot path.

A drop-in for the part of ``pulsarbat`` named in BASELINE.json: Signal containers,
``coherent_dedispersion`` / ``DispersionMeasure``, ``pb.fft.fft/ifft``, ``contrib.stft/istft``,
``to_intensity`` / ``to_stokes`` and pulse folding.  The compute lives in libpbk.so
(include/pbk.h); this package is the thin Python mirror of the reference's interface for that
path.  There is no CPU fallback: without the built library or a CUDA device, calls raise.
"""

from . import _lib  # noqa: F401
from ._lib import PbkError, PbkUnsupported  # noqa: F401
from . import units  # noqa: F401
from .units import Time  # noqa: F401
from . import core
from .core import *  # noqa: F401,F403
from . import transforms
from .transforms import *  # noqa: F401,F403
from . import pulsar
from .pulsar import *  # noqa: F401,F403
from . import contrib, fft, kernels, sharding, streaming, utils  # noqa: F401
from .device import DeviceArray  # noqa: F401

__version__ = "0.1.0"

__all__ = ["fft", "contrib", "kernels", "sharding", "streaming", "utils", "units", "Time", "DeviceArray", "PbkError",
           "PbkUnsupported"]
__all__ += core.__all__ + transforms.__all__ + pulsar.__all__
