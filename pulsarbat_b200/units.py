"""Minimal physical units and timestamps for the signal containers.

The reference leans on ``astropy.units`` / ``astropy.time`` for bookkeeping only; on the hot path
their numeric content is powers of ten between Hz and MHz, seconds, and MJD differences
(SURVEY.md 8c).  astropy is not installable in the build image, so this module provides the
small subset the containers need -- ``1 * u.MHz``, ``q.to(u.Hz)``, ``q.to_value(u.s)``,
``Time(mjd) + 3 * u.s`` -- and nothing else.  Objects coming from astropy are accepted wherever
a quantity or time is expected (they are converted through ``to_value`` / ``.mjd``).
"""

import math
import numbers

import numpy as np

__all__ = ["Unit", "Quantity", "Time", "UnitConversionError",
           "Hz", "kHz", "MHz", "GHz", "s", "ms", "us", "ns", "min", "hr", "day",
           "cycle", "one", "pc", "cm", "dm_unit", "isclose", "allclose"]


class UnitConversionError(ValueError):
    pass


def _dims_mul(a, b, sign=1):
    out = dict(a)
    for k, v in b.items():
        out[k] = out.get(k, 0) + sign * v
        if out[k] == 0:
            del out[k]
    return out


class Unit:
    """scale * product(base ** power); bases are plain strings ('s', 'cycle', 'pc', 'cm')."""

    __array_priority__ = 1000

    def __init__(self, scale, dims, name=None):
        self.scale = float(scale)
        self.dims = dict(dims)
        self.name = name

    # -- algebra ---------------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            if not other.dims and other.scale == 1.0:
                return self
            if not self.dims and self.scale == 1.0:
                return other
            return Unit(self.scale * other.scale, _dims_mul(self.dims, other.dims))
        return Quantity(other, self)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            if not other.dims and other.scale == 1.0:
                return self
            return Unit(self.scale / other.scale, _dims_mul(self.dims, other.dims, -1))
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        inv = Unit(1.0 / self.scale, {k: -v for k, v in self.dims.items()})
        return Quantity(other, inv)

    def __pow__(self, p):
        return Unit(self.scale ** p, {k: v * p for k, v in self.dims.items()})

    def __eq__(self, other):
        return isinstance(other, Unit) and self.dims == other.dims and \
            math.isclose(self.scale, other.scale, rel_tol=1e-15)

    def __hash__(self):
        return hash((round(math.log10(self.scale), 9), tuple(sorted(self.dims.items()))))

    def is_equivalent(self, other):
        return self.dims == _as_unit(other).dims

    def factor_to(self, other):
        other = _as_unit(other)
        if self.dims != other.dims:
            raise UnitConversionError(f"cannot convert '{self}' to '{other}'")
        return self.scale / other.scale

    def __repr__(self):
        if self.name:
            return self.name
        body = " ".join(f"{k}{v if v != 1 else ''}" for k, v in sorted(self.dims.items()))
        return f"{self.scale:g} {body}".strip()


def _as_unit(x):
    if isinstance(x, Unit):
        return x
    if isinstance(x, str):
        try:
            return _BY_NAME[x]
        except KeyError:
            raise UnitConversionError(f"unknown unit '{x}'")
    # astropy unit: go through its SI decomposition
    if hasattr(x, "decompose") and hasattr(x, "physical_type"):
        d = x.decompose()
        dims = {str(b): p for b, p in zip(d.bases, d.powers)}
        return Unit(float(d.scale), dims)
    raise UnitConversionError(f"not a unit: {x!r}")


class Quantity:
    """A number (or numpy array) with a unit."""

    __array_priority__ = 1000

    def __init__(self, value, unit=None):
        if isinstance(value, Quantity):
            unit = value.unit if unit is None else unit
            value = value.to_value(unit)
        self.value = value if isinstance(value, np.ndarray) else (
            np.asarray(value, dtype=float) if np.ndim(value) else float(value))
        self.unit = _as_unit(unit) if unit is not None else one

    # -- conversion ------------------------------------------------------------------------
    def to(self, unit):
        unit = _as_unit(unit)
        return Quantity(self.value * self.unit.factor_to(unit), unit)

    def to_value(self, unit=None):
        if unit is None:
            return self.value
        return self.value * self.unit.factor_to(_as_unit(unit))

    @property
    def isscalar(self):
        return np.ndim(self.value) == 0

    @property
    def shape(self):
        return np.shape(self.value)

    def __len__(self):
        return len(self.value)

    def __getitem__(self, idx):
        return Quantity(self.value[idx], self.unit)

    def __iter__(self):
        for v in self.value:
            yield Quantity(v, self.unit)

    def __float__(self):
        return float(self.to_value(one))

    # -- arithmetic ------------------------------------------------------------------------
    def _coerce(self, other):
        if isinstance(other, Quantity):
            return other
        if hasattr(other, "to_value") and hasattr(other, "unit"):  # astropy
            u_ = _as_unit(other.unit)
            return Quantity(np.asarray(other.value, dtype=float) if np.ndim(other.value)
                            else float(other.value), u_)
        return Quantity(other, one)

    def __add__(self, other):
        o = self._coerce(other)
        return Quantity(self.value + o.to_value(self.unit), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        o = self._coerce(other)
        return Quantity(self.value - o.to_value(self.unit), self.unit)

    def __rsub__(self, other):
        return self._coerce(other) - self

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit * other)
        o = self._coerce(other)
        return Quantity(self.value * o.value, self.unit * o.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit / other)
        o = self._coerce(other)
        return Quantity(self.value / o.value, self.unit / o.unit)

    def __rtruediv__(self, other):
        return self._coerce(other) / self

    def __pow__(self, p):
        return Quantity(self.value ** p, self.unit ** p)

    def _cmp(self, other, op):
        o = self._coerce(other)
        return op(self.value, o.to_value(self.unit))

    def __lt__(self, o): return self._cmp(o, np.less)
    def __le__(self, o): return self._cmp(o, np.less_equal)
    def __gt__(self, o): return self._cmp(o, np.greater)
    def __ge__(self, o): return self._cmp(o, np.greater_equal)

    def __eq__(self, o):
        try:
            return self._cmp(o, np.equal)
        except UnitConversionError:
            return False

    def __hash__(self):
        return hash((float(np.sum(self.value)), self.unit))

    def __repr__(self):
        return f"<Quantity {self.value} {self.unit}>"

    __str__ = lambda self: f"{self.value} {self.unit}"  # noqa: E731


Hz = Unit(1.0, {"s": -1}, "Hz")
kHz = Unit(1e3, {"s": -1}, "kHz")
MHz = Unit(1e6, {"s": -1}, "MHz")
GHz = Unit(1e9, {"s": -1}, "GHz")
s = Unit(1.0, {"s": 1}, "s")
ms = Unit(1e-3, {"s": 1}, "ms")
us = Unit(1e-6, {"s": 1}, "us")
ns = Unit(1e-9, {"s": 1}, "ns")
min = Unit(60.0, {"s": 1}, "min")  # noqa: A001
hr = Unit(3600.0, {"s": 1}, "hr")
day = Unit(86400.0, {"s": 1}, "day")
cycle = Unit(1.0, {"cycle": 1}, "cycle")
one = Unit(1.0, {}, "")
pc = Unit(1.0, {"pc": 1}, "pc")
cm = Unit(1.0, {"cm": 1}, "cm")
dm_unit = Unit(1.0, {"pc": 1, "cm": -3}, "pc / cm3")

_BY_NAME = {"Hz": Hz, "kHz": kHz, "MHz": MHz, "GHz": GHz, "s": s, "ms": ms, "us": us, "ns": ns,
            "min": min, "hr": hr, "day": day, "cycle": cycle, "": one, "one": one,
            "pc / cm3": dm_unit}


def to_value(q, unit):
    """Numeric value of ``q`` in ``unit`` for our quantities, astropy quantities, or (for
    dimensionless targets) plain numbers."""
    if isinstance(q, Quantity):
        return q.to_value(unit)
    if hasattr(q, "to_value"):
        return q.to_value(str(_as_unit(unit)) if not isinstance(unit, str) else unit)
    if isinstance(q, (numbers.Real, np.ndarray)):
        if _as_unit(unit).dims:
            raise UnitConversionError(f"expected a quantity in {unit}, got a bare number")
        return q
    raise UnitConversionError(f"cannot interpret {q!r} as a quantity")


def isclose(a, b, rtol=1e-9, atol=None):
    a = a if isinstance(a, Quantity) else Quantity(a)
    bv = to_value(b, a.unit) if not isinstance(b, Quantity) else b.to_value(a.unit)
    at = 0.0 if atol is None else to_value(atol, a.unit)
    return np.isclose(a.value, bv, rtol=rtol, atol=at)


def allclose(a, b, rtol=1e-9, atol=None):
    return bool(np.all(isclose(a, b, rtol=rtol, atol=atol)))


class Time:
    """UTC timestamp as (integer MJD, fractional day) -- enough for start_time bookkeeping
    (core.py:162-163 adds ``index / sample_rate``) without losing nanoseconds."""

    def __init__(self, val, val2=0.0, format="mjd", precision=9):
        if isinstance(val, Time):
            self.jd1, self.jd2 = val.jd1, val.jd2
            return
        if hasattr(val, "mjd") and hasattr(val, "jd1"):  # astropy Time
            mjd = val.jd1 - 2400000.5
            val, val2 = mjd, float(val.jd2)
        if isinstance(val, str):
            i, _, f = val.strip().partition(".")
            val, val2 = float(int(i)), float("0." + f) if f else 0.0
        if np.ndim(val) != 0 or np.ndim(val2) != 0:
            raise ValueError("Time must be scalar")
        whole = math.floor(float(val))
        frac = (float(val) - whole) + float(val2)
        carry = math.floor(frac)
        self.jd1 = whole + carry
        self.jd2 = frac - carry
        self.isscalar = True

    isscalar = True

    @property
    def mjd(self):
        return self.jd1 + self.jd2

    def __add__(self, dt):
        sec = to_value(dt, s)
        return Time(self.jd1, self.jd2 + float(sec) / 86400.0)

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, Time) or hasattr(other, "jd1"):
            other = Time(other)
            return Quantity(((self.jd1 - other.jd1) + (self.jd2 - other.jd2)) * 86400.0, s)
        return self + (-Quantity(to_value(other, s), s))

    def _key(self):
        return (self.jd1, self.jd2)

    def __lt__(self, o): return self._key() < Time(o)._key()
    def __le__(self, o): return self._key() <= Time(o)._key()
    def __gt__(self, o): return self._key() > Time(o)._key()
    def __ge__(self, o): return self._key() >= Time(o)._key()
    def __eq__(self, o): return isinstance(o, Time) and self._key() == o._key()
    def __hash__(self): return hash(self._key())

    def isclose(self, other, atol=None):
        tol = 1e-9 if atol is None else float(to_value(atol, s))
        return abs(float((self - other).to_value(s))) <= tol

    def __repr__(self):
        return f"<Time mjd={self.jd1:.0f}+{self.jd2:.15f}>"
