"""``pb.fft``: FFT functions that dispatch on the array type (reference: pulsarbat/fft.py).

The reference forwards every name in ``_FFT_FUNCS`` to ``scipy.fft``.  Only ``fft`` and ``ifft``
are on the baseband hot path (dedispersion.py:125, misc.py:47,87); those two run on the GPU for
numpy arrays and device arrays (any length: powers of two on the tile-FFT passes, other lengths
through Bluestein on top of them) and raise ``PbkUnsupported`` for what the kernels do not do
(``n=`` padding, ``norm`` other than "backward").  The other twelve
names are outside the accelerated path and are not provided: asking for them raises
AttributeError naming ``scipy.fft`` as the place to get them, rather than silently running on
the CPU under this package's name.
"""

import numpy as np

from . import kernels
from ._lib import PbkUnsupported

_GPU_FUNCS = ("fft", "ifft")
_OTHER = ("fft2", "fftn", "ifft2", "ifftn", "rfft", "rfft2", "rfftn", "irfft", "irfft2",
          "irfftn", "hfft", "ihfft")


def __dir__():
    return sorted(_GPU_FUNCS)


def _check(x, n, norm, overwrite_x, workers, plan):
    if n is not None:
        raise PbkUnsupported(-2, "n= padding/truncation is not supported by the GPU FFT")
    if norm not in (None, "backward"):
        raise PbkUnsupported(-2, f"norm={norm!r}: only the default 'backward' is supported")


def fft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    """Forward complex FFT along ``axis`` (scipy.fft.fft semantics, dtype-preserving)."""
    if n is not None and n == x.shape[axis]:
        n = None
    _check(x, n, norm, overwrite_x, workers, plan)
    if not isinstance(x, np.ndarray) and not hasattr(x, "tensor"):
        x = np.asarray(x)
    if isinstance(x, np.ndarray) and not np.iscomplexobj(x):
        x = x.astype(np.complex64 if x.dtype == np.float32 else np.complex128)
    return kernels.fft(x, axis=axis, inverse=False)


def ifft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    """Inverse complex FFT along ``axis`` with the 1/n scaling (scipy.fft.ifft semantics)."""
    if n is not None and n == x.shape[axis]:
        n = None
    _check(x, n, norm, overwrite_x, workers, plan)
    if not isinstance(x, np.ndarray) and not hasattr(x, "tensor"):
        x = np.asarray(x)
    if isinstance(x, np.ndarray) and not np.iscomplexobj(x):
        x = x.astype(np.complex64 if x.dtype == np.float32 else np.complex128)
    return kernels.fft(x, axis=axis, inverse=True)


def __getattr__(name):
    if name in _OTHER:
        raise AttributeError(
            f"pulsarbat_b200.fft.{name} is outside the accelerated baseband path; "
            f"use scipy.fft.{name} directly")
    raise AttributeError(f"module {__name__} has no attribute {name}")
