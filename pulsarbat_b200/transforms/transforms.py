"""FFT-based time and frequency shifts: GPU-backed mirror of the reference's
``transforms/transforms.py:211-361``.  Both are the dedispersion skeleton
``ifft(fft(z) * H)`` with a linear phase ramp or a band mask as H, generated inside the kernel."""

import math

import numpy as np

from .. import kernels
from .. import units as u
from ..core import BasebandSignal, Signal

__all__ = ["time_shift", "freq_shift"]


def _per_column(values, z, what):
    """Broadcast ``values`` (scalar or array matching leading sample axes, transforms.py:255-266)
    to one value per column of ``z``; returns (flat array, the UN-broadcast array with trailing
    axes of length 1).  The reference applies the phase to every column by broadcasting but walks
    the un-broadcast array when it zeroes the shifted-out part (``np.nditer`` at
    transforms.py:274 and :349), so e.g. one shift per channel of a (time, chan, pol) signal
    zeroes pol 0 only, and a scalar ``freq_shift`` zeroes column (0, 0) only -- observed by
    running the reference (tests/golden/ref_golden.npz) and reproduced here, since a drop-in
    must return what the reference returns."""
    values = np.array(values, dtype=np.float64)
    if values.ndim >= z.ndim:
        raise ValueError(f"{what} has too many dimensions. Expected <= {z.ndim - 1} dimensions, "
                         f"got {values.ndim} dimensions!")
    if values.ndim > 0:
        ix = (slice(None),) * values.ndim + (None,) * (z.ndim - values.ndim - 1)
        values = values[ix]
    try:
        full = np.broadcast_to(values, z.sample_shape)
    except ValueError as e:
        raise ValueError(f"{what} shape does not match the signal: {e}") from None
    return np.ascontiguousarray(full).reshape(-1), values


def _as_2d(z):
    data = z.data
    n = len(z)
    ncols = int(np.prod(z.sample_shape, dtype=np.int64)) if z.sample_shape else 1
    return data.reshape(n, ncols) if hasattr(data, "reshape") else np.asarray(data).reshape(n, ncols)


def time_shift(z, /, shift, crop=False):
    """Shift the signal in time by ``shift`` samples (or a time Quantity) through a phase gradient
    in the frequency domain; out-of-bounds samples are zeroed, ``crop=True`` removes them
    (drop-in for transforms.py:211-293)."""
    if not isinstance(z, Signal):
        raise TypeError("z must be a Signal.")
    if isinstance(shift, u.Quantity):
        shift = np.asarray((shift * z.sample_rate).to_value(u.one))
    flat, walked = _per_column(shift, z, "shift")
    if np.allclose(flat, 0):
        return z
    real_in = not np.iscomplexobj(np.empty(0, dtype=z.dtype))
    x2 = _as_2d(z)
    if real_in:
        x2 = np.asarray(x2).astype(np.complex128 if z.dtype == np.float64 else np.complex64)
    y = kernels.phase_ramp(x2, shift_samples=flat)
    y = np.asarray(y)
    if real_in:
        y = y.real.astype(z.dtype)
    y = np.ascontiguousarray(y).reshape(z.shape)
    start, stop = 0, 0
    it = np.nditer(walked, flags=["multi_index"])
    for a in it:
        if a < 0:
            a = int(math.floor(a))
            y[(np.s_[a:],) + it.multi_index] = 0
            stop = min(stop, a)
        else:
            a = int(math.ceil(a))
            y[(np.s_[:a],) + it.multi_index] = 0
            start = max(start, a)
    out = type(z).like(z, y)
    if crop:
        out = out[start:len(out) + stop]
    return out


def freq_shift(z, /, shift):
    """Shift the signal in frequency by mixing with a sinusoid; the part shifted out of band is
    zeroed (drop-in for transforms.py:296-361)."""
    if not isinstance(z, BasebandSignal):
        raise TypeError("Signal must be a BasebandSignal object.")
    try:
        shift_hz = np.asarray(shift.to(u.Hz).value, dtype=np.float64)
    except Exception:
        raise ValueError("shift must be a Quantity with units of frequency.") from None
    if shift_hz.ndim == 0:
        shift_hz = shift_hz[None]
    ft_flat, walked = _per_column(shift_hz / z.sample_rate_hz, z, "shift")
    n = len(z)
    x2 = _as_2d(z)
    mixed = kernels.mix(x2, ft_flat)
    # band mask [lo, hi) per column; only the columns the reference's nditer visits get one
    lo = np.zeros(z.sample_shape, np.int64)
    hi = np.zeros(z.sample_shape, np.int64)
    it = np.nditer(walked * n, flags=["multi_index"])
    for a in it:
        if a < 0:                                    # x[floor(a):] = 0   (transforms.py:352-354)
            lo[it.multi_index], hi[it.multi_index] = max(n + int(math.floor(a)), 0), n
        else:                                        # x[:ceil(a)] = 0    (transforms.py:355-357)
            lo[it.multi_index], hi[it.multi_index] = 0, min(int(math.ceil(a)), n)
    lo, hi = lo.reshape(-1), hi.reshape(-1)
    y = kernels.phase_ramp(mixed, zero_lo=lo, zero_hi=hi)
    y = y.reshape(z.shape) if hasattr(y, "reshape") else y
    return type(z).like(z, y)
