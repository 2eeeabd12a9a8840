"""Polyco-based pulse phase prediction: host-side producer of the fold kernel's phase
polynomial (reference: pulsarbat/pulsar/predictor.py).

Only what the fold path needs is mirrored: parsing tempo1 polycos (predictor.py:198-306),
evaluating phase and spin frequency (:121-147, :162-174) and re-centring the polynomial on a
block start (``phasepol``, :149-160).  Phases are returned as (integer cycles, fractional cycles)
pairs with the fraction in [-0.5, 0.5] like the reference's two-double ``Phase``
(phase.py:69-77).  The table/QTable machinery and the root finder are out of scope.
"""

from dataclasses import dataclass

import numpy as np
from numpy.polynomial import Polynomial

from .. import units as u
from ..units import Time

__all__ = ["PolycoEntry", "PhasePredictor"]


@dataclass
class PolycoEntry:
    """One polynomial span (predictor.py:17-45); ``poly`` has its domain in seconds."""

    psr: str
    obs: str
    freq: object
    tmid: Time
    span: object
    rphase: int
    poly: Polynomial

    @property
    def span_s(self):
        return float(u.to_value(self.span, u.s))


class PhasePredictor:
    """Piecewise-polynomial pulse phase predictor."""

    def __init__(self, entries):
        entries = sorted(entries, key=lambda e: (e.tmid.jd1, e.tmid.jd2))
        if not entries:
            raise ValueError("PhasePredictor needs at least one entry")
        for name in ("psr", "obs"):
            if len({getattr(e, name) for e in entries}) > 1:
                raise ValueError(f"All entries must have the same '{name}'.")
        if len({round(e.span_s, 9) for e in entries}) > 1:
            raise ValueError("All entries must have the same 'span' (Length of span).")
        if len({round(float(u.to_value(e.freq, u.Hz)), 3) for e in entries}) > 1:
            raise ValueError("All entries must have the same 'freq' (Observing frequency).")
        self.entries = entries

    def __len__(self):
        return len(self.entries)

    def __getitem__(self, idx):
        if isinstance(idx, (list, tuple, np.ndarray)):
            return PhasePredictor([self.entries[i] for i in idx])
        if isinstance(idx, slice):
            return PhasePredictor(self.entries[idx])
        return self.entries[idx]

    @property
    def intervals(self):
        """Merged validity intervals (predictor.py:85-106)."""
        iv = sorted(((e.tmid - e.span / 2, e.tmid + e.span / 2) for e in self.entries),
                    key=lambda x: x[1]._key())
        merged = []
        start, end = iv.pop()
        while iv:
            nstart, nend = iv.pop()
            if nend >= start or start.isclose(nend, 1 * u.ms):
                start = min(start, nstart, key=lambda t: t._key())
            else:
                merged.append((start, end))
                start, end = nstart, nend
        merged.append((start, end))
        return tuple(reversed(merged))

    def _index_and_dt(self, t):
        """predictor.py:108-119 for a scalar time."""
        t = Time(t)
        if not any(a <= t <= b for a, b in self.intervals):
            raise ValueError("Some timestamps outside predictor range!")
        ends = np.array([(e.tmid + e.span / 2).mjd for e in self.entries])
        idx = int(np.searchsorted(ends, t.mjd))
        idx = min(idx, len(self.entries) - 1)
        dt = float((t - self.entries[idx].tmid).to_value(u.s))
        return idx, dt

    def __call__(self, t, offsets_s=0.0):
        """(int cycles, frac cycles) at time ``t`` (+ ``offsets_s`` seconds, scalar or array)."""
        idx, dt = self._index_and_dt(t)
        e = self.entries[idx]
        val = e.poly(dt + np.asarray(offsets_s, dtype=np.float64))
        whole = np.rint(val)
        ints = (np.int64(e.rphase) + whole.astype(np.int64))
        return ints, val - whole

    def sample_phases(self, start_time, nsamp, sample_rate, *, on_device=False, device=None):
        """Phase of every sample of a block on the GPU: the reference's ``predictor(times)`` for
        ``times = start_time + arange(nsamp) / sample_rate`` (predictor.py:121-147), entry by
        entry as its ``np.unique(index)`` loop does.  Returns (int64 cycles, FP64 fraction);
        raises the reference's ValueError when a sample lies outside the predictor's range."""
        from .. import kernels
        sr = float(u.to_value(sample_rate, u.Hz))
        t0 = Time(start_time)
        self._index_and_dt(t0 + ((nsamp - 1) / sr) * u.s)          # range check of the last sample
        ends = [e.tmid + e.span / 2 for e in self.entries]
        parts, first = [], 0
        while first < nsamp:
            t_first = t0 + (first / sr) * u.s
            idx, dt = self._index_and_dt(t_first)
            left = float((ends[idx] - t_first).to_value(u.s)) * sr
            count = nsamp - first if idx == len(self.entries) - 1 else \
                max(1, min(nsamp - first, int(np.floor(left + 1e-9)) + 1))
            e = self.entries[idx]
            parts.append(kernels.predict_phase(e.poly.coef, e.rphase, nsamp=count, dt0_s=dt,
                                               sample_rate_hz=sr, on_device=on_device,
                                               device=device))
            first += count
        if len(parts) == 1:
            return parts[0]
        if on_device:
            import torch
            from ..device import DeviceArray
            return tuple(DeviceArray(torch.cat([p[k].tensor for p in parts])) for k in (0, 1))
        return tuple(np.concatenate([p[k] for p in parts]) for k in (0, 1))

    def f0(self, t, n=0):
        """Spin frequency (n=0) or its derivatives, cycles / s^(n+1) (predictor.py:162-174)."""
        idx, dt = self._index_and_dt(t)
        return float(self.entries[idx].poly.deriv(n + 1)(dt))

    def phasepol(self, t0):
        """(coefficients of the phase polynomial in seconds since ``t0`` with the integer part
        of its value at 0 removed, integer reference phase) -- predictor.py:149-160."""
        if np.ndim(getattr(t0, "mjd", 0.0)) != 0:
            raise ValueError("Timestamp must be a scalar.")
        idx, dt = self._index_and_dt(t0)
        e = self.entries[idx]
        p = e.poly.copy()
        p.domain = p.domain - dt
        a = int(p(0) // 1)
        return (p - a).convert().coef.copy(), int(e.rphase) + a

    @classmethod
    def from_polyco(cls, path):
        """Read tempo1-style polycos from a path or a file-like object (predictor.py:198-306)."""
        f = path if hasattr(path, "readline") else open(path, "r")
        d2e = str.maketrans("Dd", "ee")
        entries = []
        with f:
            while (line := f.readline()):
                if not line.strip():
                    continue
                try:
                    psr, _, _, mjd_mid, dm, *_ = line.split()
                    rphase, f0, obs, span, ncoeff, freq, *_ = f.readline().split()
                    r_int, _, r_frac = rphase.partition(".")
                    coeffs = []
                    for _ in range(-(int(ncoeff) // -3)):
                        coeffs += f.readline().translate(d2e).split()
                    coeffs = np.array(coeffs, dtype=np.float64)
                    coeffs[0] += float("0." + r_frac)
                    coeffs[1] += float(f0) * 60
                    entries.append(PolycoEntry(
                        psr=psr, obs=obs, freq=float(freq) * u.MHz, tmid=Time(mjd_mid),
                        span=int(span) * u.min, rphase=int("0" + r_int),
                        poly=Polynomial(coeffs, domain=[-60, +60]).convert()))
                except (ValueError, IndexError) as err:
                    raise ValueError(f"not a tempo1 polyco: {err}") from None
        return cls(entries)
