"""TEST INFRASTRUCTURE ONLY -- stand-in for ``astropy.time.Time`` (see the package docstring): a
(whole day, day fraction) pair of float64 scalars or arrays built from MJD numbers or MJD strings,
``t + quantity``, ``t - t`` in seconds, ordering, indexing, iteration and ``Time.isclose`` -- all
the reference's containers (core.py) and its phase predictor (pulsar/predictor.py) ask of a time.

The day fraction carries ~1e-16 day = 1e-11 s; genuine astropy keeps differences to the same
order with its two-double arithmetic, so predictor phases computed through this stand-in can
differ from astropy's by ~1e-9 cycle (the reference's own tests allow 1e-8)."""

import numpy as np

from . import units as u

__all__ = ["Time"]


def _split(val):
    """MJD number(s) or string(s) -> (whole days, fraction) as float64 arrays."""
    if isinstance(val, str):
        s = val.strip()
        if not s or s.lstrip("+-").replace(".", "", 1).isdigit() is False:
            raise ValueError(f"astropy stub: cannot read '{val}' as an MJD")
        i, _, f = s.partition(".")
        return np.float64(int(i or "0")), np.float64(float("0." + f) if f else 0.0)
    a = np.asarray(val)
    if a.dtype.kind in "US":
        parts = [_split(str(x)) for x in a.ravel()]
        return (np.array([p[0] for p in parts]).reshape(a.shape),
                np.array([p[1] for p in parts]).reshape(a.shape))
    a = a.astype(np.float64)
    whole = np.floor(a)
    return whole, a - whole


class Time:
    __array_priority__ = 200000
    __array_ufunc__ = None

    def __init__(self, val, val2=0.0, format=None, precision=9, scale="utc"):
        if isinstance(val, Time):
            self.jd1, self.jd2 = val.jd1, val.jd2
            return
        if isinstance(val, (list, tuple)) and val and isinstance(val[0], Time):
            self.jd1 = np.array([float(t.jd1) for t in val])
            self.jd2 = np.array([float(t.jd2) for t in val])
            return
        whole, frac = _split(val)
        frac = frac + np.asarray(val2, dtype=np.float64)
        carry = np.floor(frac)
        self.jd1, self.jd2 = whole + carry, frac - carry
        if np.ndim(self.jd1) == 0:
            self.jd1, self.jd2 = float(self.jd1), float(self.jd2)

    # shape protocol
    @property
    def isscalar(self):
        return np.ndim(self.jd1) == 0

    @property
    def shape(self):
        return np.shape(self.jd1)

    def __len__(self):
        return len(self.jd1)

    def __getitem__(self, ix):
        return Time._raw(np.asarray(self.jd1)[ix], np.asarray(self.jd2)[ix])

    def __iter__(self):
        for a, b in zip(self.jd1, self.jd2):
            yield Time._raw(a, b)

    @classmethod
    def _raw(cls, jd1, jd2):
        t = cls.__new__(cls)
        if np.ndim(jd1) == 0:
            jd1, jd2 = float(jd1), float(jd2)
        t.jd1, t.jd2 = jd1, jd2
        return t

    @property
    def mjd(self):
        return self.jd1 + self.jd2

    # arithmetic
    def __add__(self, dt):
        return Time(self.jd1, self.jd2 + u.Quantity._q(dt).to_value(u.s) / 86400.0)

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, Time):
            return u.Quantity(((self.jd1 - other.jd1) + (self.jd2 - other.jd2)) * 86400.0, u.s)
        return self + (-u.Quantity._q(other))

    # ordering (element-wise for arrays)
    def _diff(self, o):
        return (self.jd1 - o.jd1) + (self.jd2 - o.jd2)

    def __lt__(self, o): return self._diff(o) < 0      # noqa: E704
    def __le__(self, o): return self._diff(o) <= 0     # noqa: E704
    def __gt__(self, o): return self._diff(o) > 0      # noqa: E704
    def __ge__(self, o): return self._diff(o) >= 0     # noqa: E704

    def __eq__(self, o):
        return isinstance(o, Time) and self._diff(o) == 0

    def __hash__(self):
        return hash((float(np.sum(self.jd1)), float(np.sum(self.jd2))))

    def isclose(self, other, atol=None):
        tol = 1e-9 if atol is None else float(u.Quantity._q(atol).to_value(u.s))
        return np.abs((self - other).to_value(u.s)) <= tol

    def __repr__(self):
        return f"<Time mjd={self.jd1}+{self.jd2!r}>"
