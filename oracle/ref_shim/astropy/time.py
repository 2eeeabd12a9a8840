"""TEST INFRASTRUCTURE ONLY -- stand-in for a scalar ``astropy.time.Time`` (see the package
docstring): a (whole day, day fraction) pair, ``t + quantity``, ``t - t`` in seconds, ordering
and ``Time.isclose`` -- all the reference's containers ask of a start time."""

import math

from . import units as u

__all__ = ["Time"]


class Time:
    isscalar = True
    shape = ()

    def __init__(self, val, val2=0.0, format=None, precision=9, scale="utc"):
        if isinstance(val, Time):
            self.jd1, self.jd2 = val.jd1, val.jd2
            return
        if isinstance(val, str):
            raise ValueError("astropy stub: construct Time from MJD numbers")
        whole = math.floor(float(val))
        frac = (float(val) - whole) + float(val2)
        carry = math.floor(frac)
        self.jd1, self.jd2 = whole + carry, frac - carry

    @property
    def mjd(self):
        return self.jd1 + self.jd2

    def __add__(self, dt):
        return Time(self.jd1, self.jd2 + float(u.Quantity._q(dt).to_value(u.s)) / 86400.0)

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, Time):
            return u.Quantity(((self.jd1 - other.jd1) + (self.jd2 - other.jd2)) * 86400.0, u.s)
        return self + (-u.Quantity._q(other))

    def _key(self):
        return (self.jd1, self.jd2)

    def __lt__(self, o): return self._key() < o._key()     # noqa: E704
    def __le__(self, o): return self._key() <= o._key()    # noqa: E704
    def __gt__(self, o): return self._key() > o._key()     # noqa: E704
    def __ge__(self, o): return self._key() >= o._key()    # noqa: E704
    def __eq__(self, o): return isinstance(o, Time) and self._key() == o._key()  # noqa: E704
    def __hash__(self): return hash(self._key())           # noqa: E704

    def isclose(self, other, atol=None):
        tol = 1e-9 if atol is None else float(u.Quantity._q(atol).to_value(u.s))
        return abs(float((self - other).to_value(u.s))) <= tol

    def __repr__(self):
        return f"<Time mjd={self.jd1:.0f}+{self.jd2!r}>"
