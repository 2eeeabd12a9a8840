"""``pb.fft``: FFT functions that dispatch on the array type (reference: pulsarbat/fft.py).

The reference forwards every name in ``_FFT_FUNCS`` to ``scipy.fft`` (numpy arrays) or to
``da.fft.fft_wrap(scipy.fft.<name>)`` (dask arrays), ``fft.py:8-43``.  ``fft`` and ``ifft`` are
the two on the baseband hot path (dedispersion.py:125, misc.py:47,87): they run on the GPU for
numpy arrays, device arrays and -- chunk by chunk, lazily -- dask arrays (any length: powers of
two on the tile-FFT passes, other lengths through Bluestein), and raise ``PbkUnsupported`` for
what the kernels do not do (``n=`` padding, ``norm`` other than "backward").

The other twelve names are OUTSIDE the accelerated path.  An existing script that calls
``pb.fft.rfft`` or ``pb.fft.fft2`` must keep working after the switch, so for host and dask
arrays they are passed through to ``scipy.fft`` exactly as the reference does (same function
objects, so same results and same errors); this is not a CPU fallback inside the GPU path -- no
GPU implementation of them exists to fall back from -- and a ``DeviceArray`` argument raises
``PbkUnsupported`` rather than being copied to the host behind the caller's back.
"""

from functools import singledispatch

import numpy as np

from . import _dask, kernels
from ._lib import PbkUnsupported

_GPU_FUNCS = ("fft", "ifft")
_OTHER = ("fft2", "fftn", "ifft2", "ifftn", "rfft", "rfft2", "rfftn", "irfft", "irfft2",
          "irfftn", "hfft", "ihfft")
_FFT_FUNCS = sorted(_GPU_FUNCS + _OTHER)


def __dir__():
    return list(_FFT_FUNCS)


def _check(x, n, norm, overwrite_x, workers, plan):
    if n is not None:
        raise PbkUnsupported(-2, "n= padding/truncation is not supported by the GPU FFT")
    if norm not in (None, "backward"):
        raise PbkUnsupported(-2, f"norm={norm!r}: only the default 'backward' is supported")


def _gpu_fft(x, n, axis, norm, overwrite_x, workers, plan, inverse):
    if n is not None and n == x.shape[axis]:
        n = None
    _check(x, n, norm, overwrite_x, workers, plan)
    if _dask.is_dask(x):
        # fft.py:40-43 (da.fft.fft_wrap): the transform axis must sit in one chunk; every chunk is
        # one GPU call when the result is computed
        cdt = np.complex64 if x.dtype in (np.float32, np.complex64) else np.complex128
        return _dask.map_whole_axis(
            x, lambda b: np.asarray(_gpu_fft(np.asarray(b), None, axis, None, False, None, None,
                                             inverse)), axis, cdt)
    if not isinstance(x, np.ndarray) and not hasattr(x, "tensor"):
        x = np.asarray(x)
    if isinstance(x, np.ndarray) and not np.iscomplexobj(x):
        x = x.astype(np.complex64 if x.dtype == np.float32 else np.complex128)
    return kernels.fft(x, axis=axis, inverse=inverse)


def fft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    """Forward complex FFT along ``axis`` (scipy.fft.fft semantics, dtype-preserving)."""
    return _gpu_fft(x, n, axis, norm, overwrite_x, workers, plan, False)


def ifft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    """Inverse complex FFT along ``axis`` with the 1/n scaling (scipy.fft.ifft semantics)."""
    return _gpu_fft(x, n, axis, norm, overwrite_x, workers, plan, True)


def __getattr__(name):
    if name not in _OTHER:
        raise AttributeError(f"module {__name__} has no attribute {name}")
    import scipy.fft
    _fft_func = getattr(scipy.fft, name)

    @singledispatch
    def func(*args, **kwargs):
        if args and hasattr(args[0], "tensor"):
            raise PbkUnsupported(-2, f"pulsarbat_b200.fft.{name} has no GPU implementation; a "
                                     "DeviceArray is not copied to the host implicitly")
        return _fft_func(*args, **kwargs)

    da = _dask.dask_array()
    if da is not None:
        @func.register(da.Array)
        def _(*args, **kwargs):
            return da.fft.fft_wrap(_fft_func)(*args, **kwargs)

    func.__qualname__ = _fft_func.__qualname__
    func.__name__ = _fft_func.__name__
    func.__doc__ = _fft_func.__doc__
    return func
