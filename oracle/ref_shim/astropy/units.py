"""TEST INFRASTRUCTURE ONLY -- stand-in for ``astropy.units`` (see the package docstring).

It follows astropy's arithmetic rules for the operations the reference performs, so that the
floating-point operations executed by the reference's source are the ones astropy would issue:

* a unit is ``scale * prod(named_base ** power)``; multiplying units merges identical NAMED
  bases (MHz**2 / MHz**2 cancels exactly, Hz and MHz stay distinct until a conversion);
* ``Quantity * Quantity`` multiplies values and units, never rescales;
* ``a + b`` / ``a - b`` / comparisons convert ``b`` to ``a``'s unit (``b.value * factor``);
* ``q.to(unit)`` / ``q.to_value(unit)`` multiply the value by ONE factor
  ``decompose(self).scale / decompose(unit).scale``; a factor of exactly 1 leaves it untouched;
* ``cycle`` = 2 pi rad, as in astropy;
* ``np.fft.fftfreq(n, d=<Quantity>)`` returns ``fftfreq(n, d.value)`` in ``1 / d.unit``.

Last-ulp differences from genuine astropy in composite scale factors cannot be excluded; they
are ~1e-16 relative (1e-8 cycles on a 1e8-cycle chirp phase), far below every tolerance used."""

import math

import numpy as np

__all__ = ["Unit", "Quantity", "SpecificTypeQuantity", "UnitConversionError", "Hz", "kHz", "MHz",
           "GHz", "s", "ms", "us", "ns", "min", "day", "cycle", "rad", "one",
           "dimensionless_unscaled",
           "pc", "cm", "isclose", "allclose"]


class UnitConversionError(ValueError):
    pass


# named unit -> (SI scale, SI dims)
_NAMED = {
    "s": (1.0, {"s": 1}), "ms": (1e-3, {"s": 1}), "us": (1e-6, {"s": 1}), "ns": (1e-9, {"s": 1}),
    "min": (60.0, {"s": 1}), "day": (86400.0, {"s": 1}),
    "Hz": (1.0, {"s": -1}), "kHz": (1e3, {"s": -1}), "MHz": (1e6, {"s": -1}),
    "GHz": (1e9, {"s": -1}),
    "rad": (1.0, {"rad": 1}), "cycle": (2 * math.pi, {"rad": 1}),
    "pc": (1.0, {"pc": 1}), "cm": (1.0, {"cm": 1}),
}


def _merge(a, b, sign=1):
    out = dict(a)
    for k, v in b.items():
        p = out.get(k, 0) + sign * v
        if p == 0:
            out.pop(k, None)
        else:
            out[k] = p
    return out


class Unit:
    __array_priority__ = 100000
    __array_ufunc__ = None

    def __init__(self, scale=1.0, bases=None):
        if isinstance(scale, str):
            scale, bases = 1.0, {scale: 1}
        self.scale = scale
        self.bases = dict(bases or {})

    def decompose(self):
        """(SI scale, SI dims)."""
        scale, dims = self.scale, {}
        for name, p in self.bases.items():
            sc, d = _NAMED[name]
            scale = scale * sc ** p
            dims = _merge(dims, {k: v * p for k, v in d.items()})
        return scale, dims

    def _to(self, other):
        s0, d0 = self.decompose()
        s1, d1 = other.decompose()
        if d0 != d1:
            raise UnitConversionError(f"'{self}' and '{other}' are not convertible")
        return s0 / s1

    def to(self, other, value=1.0):
        return value * self._to(other)

    def is_equivalent(self, other):
        return self.decompose()[1] == other.decompose()[1]

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale * other.scale, _merge(self.bases, other.bases))
        if isinstance(other, Quantity):
            return Quantity(other.value, self * other.unit)
        return Quantity(other, self)

    def __rmul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit * self)
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale / other.scale, _merge(self.bases, other.bases, -1))
        if isinstance(other, Quantity):
            return Quantity(1 / other.value, self / other.unit)
        return Quantity(1 / other, self)

    def __rtruediv__(self, other):
        return Quantity(other, self ** -1)

    def __pow__(self, p):
        return Unit(self.scale ** p, {k: v * p for k, v in self.bases.items()})

    def __eq__(self, other):
        if not isinstance(other, Unit):
            return False
        a, b = self.decompose(), other.decompose()
        return a[1] == b[1] and math.isclose(a[0], b[0], rel_tol=1e-15)

    def __hash__(self):
        return hash(tuple(sorted(self.decompose()[1].items())))

    def __repr__(self):
        body = " ".join(f"{k}{p if p != 1 else ''}" for k, p in self.bases.items())
        return (f"{self.scale:g} " if self.scale != 1 else "") + body


class Quantity:
    __array_priority__ = 100000
    _default_unit = None

    def __init__(self, value, unit=None, **kw):
        if isinstance(value, Quantity):
            if unit is None:
                value, unit = value.value, value.unit
            else:
                value = value.to_value(unit)
        if unit is None:
            unit = type(self)._default_unit or one
        if isinstance(value, (np.ndarray, np.generic)):
            self.value = value if np.issubdtype(np.asarray(value).dtype, np.inexact) \
                else np.asarray(value, dtype=float)
        else:
            self.value = np.asarray(value, dtype=float) if np.ndim(value) else float(value)
        self.unit = unit

    def _new(self, value, unit):
        return Quantity(value, unit)

    # conversion
    def to(self, unit):
        return Quantity(self.to_value(unit), unit)

    def to_value(self, unit=None):
        if unit is None:
            return self.value
        scale = self.unit._to(unit)
        return self.value if scale == 1.0 else self.value * scale

    @property
    def isscalar(self):
        return np.ndim(self.value) == 0

    @property
    def shape(self):
        return np.shape(self.value)

    @property
    def ndim(self):
        return np.ndim(self.value)

    def __len__(self):
        return len(self.value)

    def __getitem__(self, ix):
        return Quantity(self.value[ix], self.unit)

    def __iter__(self):
        for v in self.value:
            yield Quantity(v, self.unit)

    def __float__(self):
        return float(self.to_value(one))

    # arithmetic
    @staticmethod
    def _q(x):
        return x if isinstance(x, Quantity) else Quantity(x, one)

    def __add__(self, o):
        return Quantity(self.value + self._q(o).to_value(self.unit), self.unit)

    def __radd__(self, o):
        return self._q(o) + self

    def __sub__(self, o):
        return Quantity(self.value - self._q(o).to_value(self.unit), self.unit)

    def __rsub__(self, o):
        return self._q(o) - self

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __abs__(self):
        return Quantity(abs(self.value), self.unit)

    def __mul__(self, o):
        if isinstance(o, Unit):
            return Quantity(self.value, self.unit * o)
        o = self._q(o)
        return Quantity(self.value * o.value, self.unit * o.unit)

    def __rmul__(self, o):
        o = self._q(o)
        return Quantity(o.value * self.value, o.unit * self.unit)

    def __truediv__(self, o):
        if isinstance(o, Unit):
            return Quantity(self.value, self.unit / o)
        o = self._q(o)
        return Quantity(self.value / o.value, self.unit / o.unit)

    def __rtruediv__(self, o):
        o = self._q(o)
        return Quantity(o.value / self.value, o.unit / self.unit)

    def __pow__(self, p):
        return Quantity(self.value ** p, self.unit ** p)

    def _cmp(self, o, op):
        if not isinstance(o, Quantity) and np.ndim(o) == 0 and (o == 0 or not np.isfinite(o)):
            return op(self.value, o)      # astropy: 0, inf and nan compare with any unit
        return op(self.value, self._q(o).to_value(self.unit))

    def __lt__(self, o): return self._cmp(o, np.less)            # noqa: E704
    def __le__(self, o): return self._cmp(o, np.less_equal)      # noqa: E704
    def __gt__(self, o): return self._cmp(o, np.greater)         # noqa: E704
    def __ge__(self, o): return self._cmp(o, np.greater_equal)   # noqa: E704

    def __eq__(self, o):
        try:
            return self._cmp(o, np.equal)
        except UnitConversionError:
            return False

    def __hash__(self):
        return hash((float(np.sum(self.value)), self.unit))

    # numpy: binary operators with ndarrays defer to the reflected methods above; the only numpy
    # function the reference calls on a quantity is np.fft.fftfreq(n, d)
    __array_ufunc__ = None

    def __array_function__(self, func, types, args, kwargs):
        if func is np.fft.fftfreq:
            n = args[0]
            d = args[1] if len(args) > 1 else kwargs["d"]
            return Quantity(np.fft.fftfreq(n, d.value), d.unit ** -1)
        if func is np.stack:
            arrs = list(args[0])
            unit = arrs[0].unit
            return Quantity(np.stack([a.to_value(unit) for a in arrs], *args[1:], **kwargs), unit)
        if func is np.unique:
            return Quantity(np.unique(args[0].value, *args[1:], **kwargs), args[0].unit)
        raise TypeError(f"astropy stub: numpy function {func.__name__} on a Quantity")

    def __repr__(self):
        return f"<Quantity {self.value} {self.unit}>"


class SpecificTypeQuantity(Quantity):
    """``_default_unit`` applies to bare numbers, as in astropy."""

    _equivalent_unit = None

    def __init__(self, value, unit=None, **kw):
        super().__init__(value, unit, **kw)
        eq = type(self)._equivalent_unit
        if eq is not None and not self.unit.is_equivalent(eq):
            raise UnitConversionError(f"{type(self).__name__} needs units equivalent to {eq}")


s, ms, us, ns, min, day = (Unit(n) for n in ("s", "ms", "us", "ns", "min", "day"))
Hz, kHz, MHz, GHz = (Unit(n) for n in ("Hz", "kHz", "MHz", "GHz"))
cycle, rad, pc, cm = (Unit(n) for n in ("cycle", "rad", "pc", "cm"))
one = dimensionless_unscaled = Unit()


def isclose(a, b, rtol=1e-05, atol=None):
    a = Quantity._q(a)
    bv = Quantity._q(b).to_value(a.unit)
    at = 0.0 if atol is None else Quantity._q(atol).to_value(a.unit)
    return np.isclose(a.value, bv, rtol=rtol, atol=at)


def allclose(a, b, rtol=1e-05, atol=None):
    return bool(np.all(isclose(a, b, rtol=rtol, atol=atol)))
