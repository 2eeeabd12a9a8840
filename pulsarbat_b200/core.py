"""Signal containers: host-side mirror of the reference's ``pulsarbat/core.py``.

Same class hierarchy, constructor signatures, validation rules, slicing/metadata behaviour and
exception types as the reference (cited per method), written against ``pulsarbat_b200.units``
instead of astropy.  The containers hold metadata only; every array operation on the hot path
(``to_intensity``, ``to_stokes``) is executed by libpbk on the GPU -- there is no numpy fallback.

Data may be a numpy array (copied to the device and back per call) or a
:class:`pulsarbat_b200.device.DeviceArray` (stays resident in HBM).
"""

import inspect
import operator

import numpy as np

from . import units as u
from .units import Time

__all__ = ["Signal", "RadioSignal", "IntensitySignal", "FullStokesSignal", "BasebandSignal",
           "DualPolarizationSignal", "InvalidSignalError"]


class InvalidSignalError(ValueError):
    """Raised for data that cannot form the requested signal type (core.py:22-25)."""


def _is_device(x):
    return type(x).__name__ == "DeviceArray"


class Signal(np.lib.mixins.NDArrayOperatorsMixin):
    """Samples along axis 0 plus a sample rate and optional start time (core.py:28-97)."""

    _req_dtype = ()
    _req_shape = (None,)
    _axes_labels = {"time": 0}

    def __init__(self, z, /, *, sample_rate, start_time=None, meta=None):
        need = len(self._req_shape)
        if z.ndim < need:
            raise InvalidSignalError(
                f"Expected signal with at least {need} dimension(s), got signal with "
                f"{z.ndim} dimension(s) instead.")
        for have, want in zip(tuple(z.shape)[:need], self._req_shape):
            if want is not None and have != want:
                raise InvalidSignalError(
                    f"Signal has invalid shape. Expected {self._req_shape}, got "
                    f"{tuple(z.shape)} instead.")
        if int(np.prod(tuple(z.shape)[1:])) == 0:
            raise InvalidSignalError("Sample shape must have non-zero size!")

        data = None
        if not self._req_dtype or z.dtype in self._req_dtype:
            data = z
        else:
            try:  # core.py:82-88: only the first allowed dtype is ever tried
                data = z.astype(self._req_dtype[0], casting="safe")
            except TypeError:
                data = None
        if data is None:
            raise InvalidSignalError(
                f"Invalid dtype. Expected {self._req_dtype}, got {z.dtype}.")

        self._data = data
        self.sample_rate = sample_rate
        self.start_time = start_time
        self.meta = meta

    # -- numpy protocol (core.py:99-122, 149-153) -------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if method != "__call__" or ufunc == np.matmul:
            return NotImplemented
        unwrap = lambda a: a.data if isinstance(a, Signal) else a  # noqa: E731
        args = tuple(unwrap(a) for a in inputs)
        outs = (None,) * ufunc.nout if out is None else out
        res = ufunc(*args, out=tuple(unwrap(o) for o in outs), **kwargs)
        if res is NotImplemented:
            return NotImplemented
        if ufunc.nout == 1:
            res = (res,)
        wrapped = tuple(type(self).like(self, r) if o is None else o for r, o in zip(res, outs))
        return wrapped[0] if len(wrapped) == 1 else wrapped

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = np.asanyarray(self.data.__array__() if _is_device(self.data) else self.data)
        return a if dtype is None else a.astype(dtype)

    def __repr__(self):
        return (f"pulsarbat_b200.{type(self).__name__}<shape={tuple(self.shape)}, "
                f"dtype={self.dtype}> @ {hex(id(self))}")

    def __str__(self):
        head = f"{type(self).__name__} @ {hex(id(self))}"
        c = type(self.data)
        lines = [head, "-" * len(head),
                 f"Data Container: {c.__module__}.{c.__name__}"
                 f"<shape={tuple(self.shape)}, dtype={self.dtype}>"]
        lines += self._attr_lines()
        return "\n".join(lines)

    def _attr_lines(self):
        st = "N/A" if self.start_time is None else repr(self.start_time)
        return [f"Sample rate: {self.sample_rate}", f"Time length: {self.time_length}",
                f"Start time: {st}"]

    # -- slicing (core.py:155-176) ----------------------------------------------------------
    def _time_slice(self, index):
        sl = slice(*index.indices(self.shape[0]))
        assert sl.step > 0, "Time axis slicing does not support negative step"
        kw = {}
        if sl.step > 1:
            kw["sample_rate"] = self.sample_rate / sl.step
        if self.start_time is not None:
            kw["start_time"] = self.start_time + sl.start / self.sample_rate
        return kw

    def _slice_kwargs(self, index):
        if not all(isinstance(a, slice) for a in index[:1]):
            raise IndexError("Only supports slicing on time axis.")
        return self._time_slice(index[0])

    def __getitem__(self, index):
        if not isinstance(index, tuple):
            index = (index,)
        kw = self._slice_kwargs(index)
        return type(self).like(self, self.data[index], **kw)

    def get_axis(self, axis):
        try:
            axis = operator.index(axis)
        except TypeError:
            axis = self.axes_labels.get(axis, None)
        if axis is None or axis < -self.ndim or self.ndim <= axis:
            raise ValueError("Invalid axis.")
        return axis

    # -- properties -------------------------------------------------------------------------
    @property
    def axes_labels(self):
        return self._axes_labels

    @property
    def meta(self):
        return self._meta

    @meta.setter
    def meta(self, meta):
        try:
            self._meta = None if meta is None else dict(meta)
        except Exception:
            raise ValueError("meta must be a dict.")

    @property
    def data(self):
        return self._data

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def sample_shape(self):
        return self.shape[1:]

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def sample_rate(self):
        return self._sample_rate

    @sample_rate.setter
    def sample_rate(self, sample_rate):
        try:
            hz = u.to_value(sample_rate, u.Hz)
            assert np.ndim(hz) == 0 and hz > 0
        except Exception:
            raise ValueError("Invalid sample_rate. Must be a positive scalar Quantity with "
                             "units of Hz or equivalent.")
        self._sample_rate = sample_rate
        self._sample_rate_hz = float(hz)

    @property
    def sample_rate_hz(self):
        """Sample rate as a plain float in Hz (what the kernels consume)."""
        return self._sample_rate_hz

    @property
    def dt(self):
        return (1 / self.sample_rate).to(u.s)

    @property
    def time_length(self):
        return (len(self) / self.sample_rate).to(u.s)

    @property
    def start_time(self):
        return self._start_time

    @start_time.setter
    def start_time(self, start_time):
        try:
            self._start_time = None if start_time is None else Time(start_time)
        except Exception:
            raise ValueError("Invalid start_time. Must be a scalar Time object.")

    @property
    def stop_time(self):
        if self.start_time is None:
            return None
        return self.start_time + self.time_length

    def contains(self, t, /):
        """Whether a time is within [start, stop) (core.py:285-292)."""
        if self.start_time is None:
            return False
        t = Time(t)
        t0, t1 = self.start_time, self.stop_time
        edge = (not t.isclose(t1)) or t.isclose(t0)
        return bool(edge and t0 <= t and t < t1)

    __contains__ = contains

    # -- data movement ----------------------------------------------------------------------
    def compute(self, **kwargs):
        """Signal with host (numpy) data; device data is copied back (core.py:298-309)."""
        return type(self).like(self, np.asarray(self))

    persist = compute

    def to_device(self, device=0):
        """Signal with the same metadata and data resident on the GPU."""
        from .device import DeviceArray
        if _is_device(self.data):
            return self
        return type(self).like(self, DeviceArray.from_numpy(np.ascontiguousarray(self.data),
                                                            device))

    @classmethod
    def like(cls, obj, z=None, /, **kwargs):
        """New signal of this class taking unspecified arguments from ``obj`` (core.py:347-379)."""
        for name, par in inspect.signature(cls).parameters.items():
            if par.kind is par.POSITIONAL_ONLY or name in kwargs:
                continue
            if hasattr(obj, name):
                kwargs[name] = getattr(obj, name)
            elif par.default is par.empty:
                raise ValueError(f"Missing required keyword argument: {name}")
        return cls(obj.data if z is None else z, **kwargs)


class RadioSignal(Signal):
    """Heterodyned multi-channel signal, shape (nsample, nchan, ...) (core.py:382-574)."""

    _req_shape = (None, None)
    _axes_labels = {"time": 0, "freq": 1}

    def __init__(self, z, /, *, sample_rate, start_time=None, center_freq, chan_bw,
                 freq_align="center", meta=None):
        super().__init__(z, sample_rate=sample_rate, start_time=start_time, meta=meta)
        self.center_freq = center_freq
        self.chan_bw = chan_bw
        self.freq_align = freq_align

    def _attr_lines(self):
        return super()._attr_lines() + [f"Channel Bandwidth: {self.chan_bw}",
                                        f"Total Bandwidth: {self.bandwidth}",
                                        f"Center Frequency: {self.center_freq}"]

    def _freq_slice(self, index):
        sl = slice(*index.indices(self.shape[1]))
        assert sl.step == 1, "Does not support slice step for frequency axis"
        assert sl.stop > sl.start, "Empty frequency slice!"
        f = self.channel_freqs[sl]
        return {"center_freq": (f[0] + f[-1]) / 2, "freq_align": "center"}

    def _slice_kwargs(self, index):
        if not all(isinstance(a, slice) for a in index[:2]):
            raise IndexError("Only supports slicing on time and frequency axes.")
        kw = self._time_slice(index[0])
        if len(index) > 1:
            kw.update(self._freq_slice(index[1]))
        return kw

    @property
    def nchan(self):
        return self.shape[1]

    @property
    def center_freq(self):
        return self._center_freq

    @center_freq.setter
    def center_freq(self, center_freq):
        try:
            hz = u.to_value(center_freq, u.Hz)
            assert np.ndim(hz) == 0
        except Exception:
            raise ValueError("Invalid center_freq. Must be a scalar Quantity with units of Hz "
                             "or equivalent.")
        self._center_freq = center_freq
        self._center_freq_hz = float(hz)

    @property
    def chan_bw(self):
        return self._chan_bw

    @chan_bw.setter
    def chan_bw(self, chan_bw):
        try:
            hz = u.to_value(chan_bw, u.Hz)
            assert np.ndim(hz) == 0 and hz > 0
        except Exception:
            raise ValueError("Invalid chan_bw. Must be a positive scalar Quantity with units "
                             "of Hz or equivalent.")
        self._chan_bw = chan_bw
        self._chan_bw_hz = float(hz)

    @property
    def bandwidth(self):
        return self.chan_bw * self.nchan

    @property
    def max_freq(self):
        return self.center_freq + self.bandwidth / 2

    @property
    def min_freq(self):
        return self.center_freq - self.bandwidth / 2

    @property
    def freq_align(self):
        return self._freq_align

    @freq_align.setter
    def freq_align(self, freq_align):
        if freq_align not in {"bottom", "center", "top"}:
            raise ValueError("Invalid freq_align. Expected: {'bottom', 'center', 'top'}")
        self._freq_align = "center" if self.nchan % 2 else freq_align

    @property
    def channel_freqs(self):
        """Channel centre frequencies (core.py:569-574)."""
        a = {"bottom": 0, "center": 0.5, "top": 1}[self.freq_align]
        ids = np.arange(self.nchan) + a - self.nchan / 2
        return self.center_freq + self.chan_bw * ids

    @property
    def channel_freqs_hz(self):
        """Same numbers as ``channel_freqs`` as a float64 array in Hz."""
        a = {"bottom": 0, "center": 0.5, "top": 1}[self.freq_align]
        ids = np.arange(self.nchan) + a - self.nchan / 2
        return self._center_freq_hz + self._chan_bw_hz * ids


class IntensitySignal(RadioSignal):
    """Real-valued power (core.py:577-611)."""

    _req_dtype = (np.float64, np.float32)


class FullStokesSignal(IntensitySignal):
    """(nsample, nchan, 4) Stokes [I, Q, U, V] (core.py:614-701)."""

    _req_shape = (None, None, 4)
    _axes_labels = {"time": 0, "freq": 1, "pol": 2}
    _stokes_ids = {"I": 0, "Q": 1, "U": 2, "V": 3}

    def __getitem__(self, key):
        if not isinstance(key, str):
            return super().__getitem__(key)
        if key not in self._stokes_ids:
            raise KeyError("Invalid key. Should be in {'I', 'Q', 'U', 'V'}.")
        x = np.take(np.asarray(self), self._stokes_ids[key], axis=self.get_axis("pol"))
        return IntensitySignal.like(self, x)

    stokesI = property(lambda self: self["I"])
    stokesQ = property(lambda self: self["Q"])
    stokesU = property(lambda self: self["U"])
    stokesV = property(lambda self: self["V"])


class BasebandSignal(RadioSignal):
    """Nyquist-sampled complex baseband: chan_bw == sample_rate (core.py:704-774)."""

    _req_dtype = (np.complex128, np.complex64)

    def __init__(self, z, /, *, sample_rate, start_time=None, center_freq, freq_align="center",
                 meta=None):
        super().__init__(z, sample_rate=sample_rate, start_time=start_time,
                         center_freq=center_freq, chan_bw=sample_rate, freq_align=freq_align,
                         meta=meta)

    def to_intensity(self):
        """re**2 + im**2 per element (core.py:766-774), computed on the GPU."""
        from . import _dask, kernels
        if _dask.is_dask(self.data):
            return IntensitySignal.like(self, _dask.map_elementwise(
                self.data, lambda b: np.asarray(kernels.detect(np.asarray(b), stokes=False)),
                out_dtype=_dask.real_dtype_of(self.dtype)))
        return IntensitySignal.like(self, kernels.detect(self.data, stokes=False))


class DualPolarizationSignal(BasebandSignal):
    """(nsample, nchan, 2) dual-polarisation baseband (core.py:777-966)."""

    _req_shape = (None, None, 2)
    _axes_labels = {"time": 0, "freq": 1, "pol": 2}

    def __init__(self, z, /, *, sample_rate, start_time=None, center_freq, freq_align="center",
                 pol_type, meta=None):
        super().__init__(z, sample_rate=sample_rate, start_time=start_time,
                         center_freq=center_freq, freq_align=freq_align, meta=meta)
        self.pol_type = pol_type

    def _attr_lines(self):
        basis = {"linear": "[X, Y]", "circular": "[L, R]"}[self.pol_type]
        return super()._attr_lines() + [f"Polarization Type: {self.pol_type} {basis}"]

    @property
    def pol_type(self):
        return self._pol_type

    @pol_type.setter
    def pol_type(self, pol_type):
        if pol_type not in {"linear", "circular"}:
            raise ValueError("pol_type must be in {'linear', 'circular'}")
        self._pol_type = pol_type

    def to_linear(self):
        """[L, R] -> [X, Y] = [L + R, i (L - R)] / sqrt(2) (core.py:882-904)."""
        from . import kernels
        if self.pol_type != "circular":
            return type(self).like(self, self.data, pol_type="linear")
        return type(self).like(self, kernels.pol_basis(self.data, to_circular=False),
                               pol_type="linear")

    def to_circular(self):
        """[X, Y] -> [L, R] = [X - iY, X + iY] / sqrt(2) (core.py:906-928)."""
        from . import kernels
        if self.pol_type != "linear":
            return type(self).like(self, self.data, pol_type="circular")
        return type(self).like(self, kernels.pol_basis(self.data, to_circular=True),
                               pol_type="circular")

    def to_stokes(self):
        """IQUV in the PSR/IEEE convention (core.py:930-966), computed on the GPU."""
        from . import _dask, kernels
        if _dask.is_dask(self.data):
            x, pt = self.data.rechunk({2: -1}), self.pol_type
            return FullStokesSignal.like(self, _dask.map_elementwise(
                x, lambda b: np.asarray(kernels.stokes(np.asarray(b), pt)),
                out_dtype=_dask.real_dtype_of(self.dtype), chunks=x.chunks[:2] + ((4,),)))
        return FullStokesSignal.like(self, kernels.stokes(self.data, self.pol_type))

    def to_stokes_I(self):
        """Stokes I only (AA + BB, identical in both bases, core.py:948/960)."""
        from . import _dask, kernels
        if _dask.is_dask(self.data):
            x = self.data.rechunk({2: -1})
            return IntensitySignal.like(self, _dask.map_elementwise(
                x, lambda b: np.asarray(kernels.detect(np.asarray(b), stokes=True)),
                out_dtype=_dask.real_dtype_of(self.dtype), chunks=x.chunks[:2], drop_axis=2))
        return IntensitySignal.like(self, kernels.detect(self.data, stokes=True))
