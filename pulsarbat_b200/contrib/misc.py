"""Critically-sampled STFT / inverse ("channelize / unchannelize"), GPU-backed mirror of the
reference's ``contrib/misc.py``."""

import numpy as np

from .. import _dask, kernels
from ..core import BasebandSignal

__all__ = ["stft", "istft"]


def stft(z, /, window="boxcar", nperseg=256, noverlap=0, nfft=None):
    """Short-time Fourier transform with a boxcar window and no overlap (misc.py:17-55).

    Same contract as the reference: unsupported window/overlap/nfft returns ``NotImplemented``
    (misc.py:31-32), a non-baseband signal raises ``ValueError`` (misc.py:34-35); the result has
    ``sample_rate / nperseg`` and ``freq_align`` 'bottom' for even nperseg, 'center' for odd.
    """
    if window != "boxcar" or noverlap != 0 or nfft is not None:
        return NotImplemented
    if not isinstance(z, BasebandSignal):
        raise ValueError("z must be a BasebandSignal.")
    n = int(nperseg)
    # the reference slices BOTH axes (misc.py:41): the frequency slice re-centres center_freq on
    # the mean of the channel centres and sets freq_align='center' (core.py:479-484) before
    # ``like`` applies the new freq_align, so for 'bottom'/'top' inputs with an even number of
    # channels the output's center_freq is shifted by half a coarse channel exactly as there
    z = z[: len(z) - len(z) % n, :]
    if _dask.is_dask(z.data):      # per time chunk (whole segments), lazily
        x = _dask.map_time_chunks(z.data, lambda b: np.asarray(kernels.stft(np.asarray(b), n)),
                                  multiple=n, rows_out=lambda m: m // n, cols_out=z.nchan * n,
                                  out_dtype=z.dtype)
    else:
        x = kernels.stft(z.data, n)
    falign = "center" if n % 2 else "bottom"
    return type(z).like(z, x, sample_rate=z.sample_rate / n, freq_align=falign)


def istft(z, /, window="boxcar", nperseg=256, noverlap=0, nfft=None):
    """Inverse of :func:`stft` (misc.py:58-93).  Unlike the reference (misc.py:82-83) the input
    array is not modified."""
    if window != "boxcar" or noverlap != 0 or nfft is not None:
        return NotImplemented
    if not isinstance(z, BasebandSignal):
        raise ValueError("z must be a BasebandSignal.")
    n = int(nperseg)
    if _dask.is_dask(z.data):
        x = _dask.map_time_chunks(z.data, lambda b: np.asarray(kernels.istft(np.asarray(b), n)),
                                  multiple=1, rows_out=lambda m: m * n, cols_out=z.nchan // n,
                                  out_dtype=z.dtype)
    else:
        x = kernels.istft(z.data, n)
    return type(z).like(z, x, sample_rate=z.sample_rate * n, freq_align="center")
