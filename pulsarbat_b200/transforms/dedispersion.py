"""Dedispersion: GPU-backed mirror of the reference's ``transforms/dedispersion.py``.

``coherent_dedispersion`` keeps the reference signature, return type, metadata update and
exceptions (dedispersion.py:81-133); the FFT, the chirp and the inverse FFT run as fused
sm_100a passes in libpbk (the chirp is generated in registers and never stored).
"""

import math

import numpy as np

from .. import _dask
from .. import _lib as L
from .. import kernels
from .. import units as u
from ..core import BasebandSignal, IntensitySignal, RadioSignal

__all__ = ["DispersionMeasure", "DM", "coherent_dedispersion", "incoherent_dedispersion",
           "dedisperse_detect", "overlap_save_dedispersion"]

#: s MHz^2 cm^3 / pc -- dedispersion.py:30
DISPERSION_CONSTANT = 1.0 / 2.41e-4


def _hz(q):
    """Frequency in Hz; bare infinities are allowed as reference frequencies
    (reference tests/test_dedispersion.py:16-19 pass ``np.inf``)."""
    if isinstance(q, (int, float, np.floating)) and math.isinf(q):
        return float(q)
    return float(u.to_value(q, u.Hz))


class DispersionMeasure(u.Quantity):
    """Dispersion measure in pc / cm^3 (dedispersion.py:26-75)."""

    dispersion_constant = DISPERSION_CONSTANT

    def __init__(self, value, unit=None):
        if isinstance(value, u.Quantity):
            value = value.to_value(u.dm_unit)
        elif hasattr(value, "to_value"):
            value = float(value.to_value("pc / cm3"))
        super().__init__(value, u.dm_unit)

    def __neg__(self):
        return DispersionMeasure(-self.value)

    def __mul__(self, other):
        r = super().__mul__(other)
        return DispersionMeasure(r.value) if r.unit == u.dm_unit else r

    __rmul__ = __mul__

    @property
    def dm(self):
        return float(self.value)

    def time_delay(self, f, ref_freq):
        """Delay of frequency ``f`` relative to ``ref_freq`` (dedispersion.py:32-36)."""
        f_mhz = np.asarray(_hz_array(f)) / 1e6
        r_mhz = _hz(ref_freq) / 1e6
        with np.errstate(divide="ignore"):
            delay = self.dispersion_constant * self.value * (1 / f_mhz ** 2 - 1 / r_mhz ** 2)
        return u.Quantity(delay, u.s)

    def sample_delay(self, f, ref_freq, sample_rate):
        """Delay in (fractional) samples (dedispersion.py:38-42)."""
        d = self.time_delay(f, ref_freq).to_value(u.s) * float(u.to_value(sample_rate, u.Hz))
        return float(d) if np.ndim(d) == 0 else d

    def chirp_function(self, N, dt, center_freq, ref_freq):
        """(N,) complex64 transfer function for one channel (dedispersion.py:44-57)."""
        sr = 1.0 / float(u.to_value(dt, u.s))
        return kernels.chirp(int(N), 1, dm=self.dm, sample_rate_hz=sr,
                             ref_freq_hz=_hz(ref_freq), chan_freq_hz=[_hz(center_freq)])[:, 0]

    def chirp_from_signal(self, z, /, *, ref_freq=None):
        """(N, nchan, 1, ...) complex64 chirp for ``z`` (dedispersion.py:59-75)."""
        if not isinstance(z, BasebandSignal):
            raise TypeError("Signal must be a BasebandSignal object.")
        if ref_freq is None:
            ref_freq = z.center_freq
        c = kernels.chirp(len(z), z.nchan, dm=self.dm, sample_rate_hz=z.sample_rate_hz,
                          ref_freq_hz=_hz(ref_freq), chan_freq_hz=z.channel_freqs_hz)
        return c.reshape(c.shape + (1,) * (z.ndim - 2))


DM = DispersionMeasure


def _hz_array(f):
    if isinstance(f, (int, float, np.floating, np.ndarray)) and np.all(np.isinf(f)):
        return np.asarray(f, dtype=float)
    return np.asarray(u.to_value(f, u.Hz), dtype=float)


def _as_dm(x):
    return x if isinstance(x, DispersionMeasure) else DispersionMeasure(x)


def crop_range(z, dm, ref_freq):
    """(start, stop) exactly as dedispersion.py:127-131."""
    d_top = dm.sample_delay(z.max_freq, ref_freq, z.sample_rate)
    d_bot = dm.sample_delay(z.min_freq, ref_freq, z.sample_rate)
    start = math.ceil(-min(0, d_top, d_bot))
    stop = len(z) - math.ceil(+max(0, d_top, d_bot))
    return start, stop


def _cropped_like(cls, z, data, start, **kw):
    """``cls.like(z, x)[start:stop]`` for data that is already cropped: same start_time update as
    core.py:162-163."""
    if z.start_time is not None:
        kw.setdefault("start_time", z.start_time + start / z.sample_rate)
    return cls.like(z, data, **kw)


def coherent_dedispersion(z, DM, /, *, ref_freq=None, chirp=None):
    """Coherently dedisperse a baseband signal (drop-in for dedispersion.py:81-133).

    Returns ``type(z)`` cropped at both ends by the dispersion sweep; raises ``TypeError`` for
    non-baseband input.  With ``chirp=`` the given array multiplies the spectrum instead of the
    generated chirp (dedispersion.py:121-124).  Any length works (powers of two are the fast
    path); there is no CPU fallback.
    """
    if not isinstance(z, BasebandSignal):
        raise TypeError("Signal must be a BasebandSignal object.")
    DM = _as_dm(DM)
    if ref_freq is None:
        ref_freq = z.center_freq
    start, stop = crop_range(z, DM, ref_freq)
    if _dask.is_dask(z.data):
        # lazy input: one GPU call per channel chunk when the result is computed, every chunk
        # with ITS channel frequencies and the GLOBAL ref_freq and crop (transforms.py:49-50)
        freqs, sr, rf, dm = z.channel_freqs_hz, z.sample_rate_hz, _hz(ref_freq), DM.dm
        ch = None if chirp is None else np.asarray(chirp).reshape(len(z), z.nchan)

        def chunk(block, lo, hi):
            return np.asarray(kernels.dedisperse(
                np.asarray(block), dm=dm, sample_rate_hz=sr, chan_freq_hz=freqs[lo:hi],
                ref_freq_hz=rf, crop=(start, stop),
                chirp_array=None if ch is None else ch[:, lo:hi]))
        x = _dask.map_channel_chunks(z.data, chunk, out_rows=max(0, stop - start),
                                     out_dtype=z.dtype)
    else:
        x = kernels.dedisperse(z.data, dm=DM.dm, sample_rate_hz=z.sample_rate_hz,
                               chan_freq_hz=z.channel_freqs_hz, ref_freq_hz=_hz(ref_freq),
                               crop=(start, stop), chirp_array=chirp)
    if stop <= start:
        # empty crop: the reference would hand back a zero-length slice (dedispersion.py:133)
        start = min(start, len(z))
    return _cropped_like(type(z), z, x, start)


def incoherent_delays(z, DM, ref_freq):
    """(delays, N_out, crop_before) exactly as dedispersion.py:164-169."""
    delays = np.asarray(DM.sample_delay(z.channel_freqs, ref_freq, z.sample_rate), dtype=float)
    delays = np.atleast_1d(delays).round().astype(np.int64)
    crop_before = -min(0, int(delays[0]), int(delays[-1]))
    delays = delays + crop_before
    return delays, len(z) - int(max(delays)), crop_before


def incoherent_dedispersion(z, DM, /, *, ref_freq=None):
    """Incoherently dedisperse a signal: per-channel integer roll and crop (drop-in for
    dedispersion.py:136-177; same TypeError, crop and start_time update).  Any real or complex
    RadioSignal; the roll is a bit-exact gather on the GPU."""
    if not isinstance(z, RadioSignal):
        raise TypeError("Signal must be a RadioSignal object.")
    DM = _as_dm(DM)
    if ref_freq is None:
        ref_freq = z.center_freq
    delays, n_out, crop_before = incoherent_delays(z, DM, ref_freq)
    if n_out < 0 or int(delays.min()) < 0:
        # numpy's negative slices would wrap here (sweep longer than the signal); refuse instead
        raise ValueError("dispersion sweep exceeds the signal length")
    x = kernels.shift_channels(z.data, delays, n_out)
    kw = {}
    if crop_before and z.start_time is not None:
        kw["start_time"] = z.start_time + crop_before / z.sample_rate
    return type(z).like(z, x, **kw)


def dedisperse_detect(z, DM, /, *, ref_freq=None, stokes_I=False, downsample=1, crop=True):
    """Fused coherent dedispersion -> power detection -> time sum in one plan.

    Equivalent to ``coherent_dedispersion(z, DM).to_intensity()`` (or ``to_stokes_I``) followed by
    summing ``downsample`` consecutive samples, without materialising the dedispersed voltages.
    ``crop=False`` keeps the whole circular result (for configurations whose sweep exceeds the
    block).  Accepts int8 (re, im) pairs in a trailing axis of length 2 when
    ``z`` is given as a tuple ``(raw_int8, template_signal)``.
    """
    raw = None
    if isinstance(z, tuple):
        raw, z = z
    if not isinstance(z, BasebandSignal):
        raise TypeError("Signal must be a BasebandSignal object.")
    DM = _as_dm(DM)
    if ref_freq is None:
        ref_freq = z.center_freq
    start, stop = crop_range(z, DM, ref_freq) if crop else (0, len(z))
    kind = L.OUT_STOKES_I if stokes_I else L.OUT_INTENSITY
    src = z.data if raw is None else raw
    if _dask.is_dask(src):
        freqs, sr, rf, dm = z.channel_freqs_hz, z.sample_rate_hz, _hz(ref_freq), DM.dm

        def chunk(block, lo, hi):
            return np.asarray(kernels.dedisperse(
                np.asarray(block), dm=dm, sample_rate_hz=sr, chan_freq_hz=freqs[lo:hi],
                ref_freq_hz=rf, crop=(start, stop), out_kind=kind, downsample=downsample,
                int8=raw is not None))
        rows = max(0, stop - start) // int(downsample)
        if raw is not None:        # (N, C, P, 2) int8: the (re, im) axis never survives
            def chunk_raw(block, lo, hi):
                y = chunk(block, lo, hi)
                return y.reshape(y.shape + (1,) * (block.ndim - y.ndim))
            x = _dask.map_channel_chunks(src, chunk_raw, out_rows=rows, out_dtype=np.float32)
            x = x.reshape(x.shape[:2] + (() if stokes_I else tuple(z.shape[2:])))
        else:
            x = _dask.map_channel_chunks(src, chunk, out_rows=rows, drop_trailing=stokes_I,
                                         out_dtype=_dask.real_dtype_of(z.dtype))
    else:
        x = kernels.dedisperse(src, dm=DM.dm, sample_rate_hz=z.sample_rate_hz,
                               chan_freq_hz=z.channel_freqs_hz, ref_freq_hz=_hz(ref_freq),
                               crop=(start, stop), out_kind=kind, downsample=downsample,
                               int8=raw is not None)
    kw = {"chan_bw": z.chan_bw}
    if downsample > 1:
        kw["sample_rate"] = z.sample_rate / downsample
    return _cropped_like(IntensitySignal, z, x, max(start, 0), **kw)


def overlap_save_dedispersion(z, DM, block_len, /, *, ref_freq=None):
    """Dedisperse a long stream in overlapping blocks of ``block_len`` samples (builder-defined,
    SURVEY.md 8a row O): the result equals the concatenation of ``coherent_dedispersion`` applied
    to blocks that advance by the valid length, i.e. what the reference gives per block."""
    if not isinstance(z, BasebandSignal):
        raise TypeError("Signal must be a BasebandSignal object.")
    DM = _as_dm(DM)
    if ref_freq is None:
        ref_freq = z.center_freq
    probe = z[:block_len]
    start, stop = crop_range(probe, DM, ref_freq)
    valid = stop - start
    if valid <= 0:
        raise ValueError("block length does not exceed the dispersion sweep")
    starts = list(range(0, len(z) - block_len + 1, valid))
    kw = dict(dm=DM.dm, sample_rate_hz=z.sample_rate_hz, chan_freq_hz=z.channel_freqs_hz,
              ref_freq_hz=_hz(ref_freq), crop=(start, stop))
    if isinstance(z.data, kernels.DeviceArray) or np.asarray(z.data).dtype != np.complex64:
        pieces = [np.asarray(kernels.dedisperse(z[b:b + block_len].data, **kw)) for b in starts]
        data = np.concatenate(pieces, axis=0)
    else:
        # host complex64 stream: the blocks go through the three-stream pipeline (upload of block
        # i+1, kernels of block i and download of block i-1 overlap) straight into the result
        from .. import streaming
        src = np.asarray(z.data)
        data = np.empty((len(starts) * valid,) + src.shape[1:], np.complex64)
        blocks = (src[b:b + block_len] for b in starts)
        for i, y in enumerate(streaming.dedisperse_blocks(blocks, pinned_out=True, **kw)):
            data[i * valid:(i + 1) * valid] = y
    return _cropped_like(type(z), z, data, start)
