"""GPU-backed mirror of the FFT-based part of the reference's ``pulsarbat/utils.py``
(SURVEY.md 8f rank 3): ``real_to_complex``, the ingest step for real-sampled baseband
(reference readers/_baseband_readers.py:140-142)."""

import numpy as np

from . import _lib as L
from . import kernels
from .device import DeviceArray

__all__ = ["real_to_complex"]


def real_to_complex(z, axis=0):
    """Complex baseband representation of a real signal (drop-in for utils.py:15-65): analytic
    signal through a Hilbert mask in the frequency domain, shift by -B/2, decimate by 2.

    On the GPU this is ONE phase-ramp plan (real float32 in, mask generated in the middle pass)
    followed by a sign-and-decimate kernel: exp(-i pi n / 2) at the kept samples n = 2m is
    exactly (-1)^m.  float64 input is computed in float32 and returned as complex128.
    """
    dev_in = isinstance(z, DeviceArray)
    if not dev_in:
        z = np.asarray(z)
    if np.iscomplexobj(np.empty(0, dtype=z.dtype)):
        raise ValueError("Input must be real-valued.")
    out_dtype = np.complex64 if z.dtype == np.float32 else np.complex128
    n = z.shape[axis]
    if n == 0:
        return z.astype(out_dtype)
    if dev_in:
        if axis % z.ndim != 0:
            raise L.PbkUnsupported(-2, "device arrays are converted along axis 0 only")
        x = z.contiguous()
        if x.dtype != np.float32:
            x = x.astype(np.float32)
        shape = x.shape
        x2 = DeviceArray(x.tensor.reshape(n, -1))
    else:
        x = np.moveaxis(z, axis, 0)
        shape = x.shape
        x2 = np.ascontiguousarray(x.reshape(n, -1), dtype=np.float32)
    ncols = x2.shape[1]
    y = kernels.analytic_decimate(x2)
    out_shape = ((n + 1) // 2,) + tuple(shape[1:])
    if dev_in:
        return DeviceArray(y.tensor.reshape(out_shape))
    y = y.reshape(out_shape)
    return np.moveaxis(y, 0, axis).astype(out_dtype, copy=False) if ncols else y
