"""A small, self-contained unit algebra for processing-chain configs.

The reference delegates to ``pint`` (src/dspeed/units.py:1-6; uses at
processing_chain.py:63-64, 96-124, 841-845, 1749-1766).  The chain only ever needs
time and frequency units (``ns us ms s``, ``Hz kHz MHz GHz``), their products,
quotients and powers, and a way to turn "a quantity x the grid period to some power"
into a pure number (what the reference obtains from ``pi_theorem``).  Everything else
(``ADC``, ``ADC/sample``, ``mV`` ...) is an opaque label, exactly as in the reference,
where such strings are simply not found in the registry.

When ``pint`` is installed, ``pint.Quantity`` / ``pint.Unit`` objects handed to the chain
are converted with :func:`from_foreign`.
"""

from __future__ import annotations

from fractions import Fraction
from numbers import Real

# name -> (scale to seconds**dim, dim)   dim = exponent of time.  Scales are kept as exact
# rationals so that e.g. 1500 samples * 16 ns is exactly 24000 ns and 10 us / 16 ns is exactly 625.
_UNITS = {
    "s": (Fraction(1), 1), "second": (Fraction(1), 1), "seconds": (Fraction(1), 1), "sec": (Fraction(1), 1),
    "ms": (Fraction(1, 10**3), 1), "millisecond": (Fraction(1, 10**3), 1), "milliseconds": (Fraction(1, 10**3), 1),
    "us": (Fraction(1, 10**6), 1), "µs": (Fraction(1, 10**6), 1), "microsecond": (Fraction(1, 10**6), 1),
    "microseconds": (Fraction(1, 10**6), 1),
    "ns": (Fraction(1, 10**9), 1), "nanosecond": (Fraction(1, 10**9), 1), "nanoseconds": (Fraction(1, 10**9), 1),
    "ps": (Fraction(1, 10**12), 1), "picosecond": (Fraction(1, 10**12), 1),
    "min": (Fraction(60), 1), "minute": (Fraction(60), 1), "hour": (Fraction(3600), 1), "h": (Fraction(3600), 1),
    "Hz": (Fraction(1), -1), "hertz": (Fraction(1), -1),
    "kHz": (Fraction(10**3), -1), "kilohertz": (Fraction(10**3), -1),
    "MHz": (Fraction(10**6), -1), "megahertz": (Fraction(10**6), -1),
    "GHz": (Fraction(10**9), -1), "gigahertz": (Fraction(10**9), -1),
    "dimensionless": (Fraction(1), 0),
}
_SYMBOL = {
    "second": "s", "seconds": "s", "sec": "s", "millisecond": "ms", "milliseconds": "ms",
    "microsecond": "us", "microseconds": "us", "µs": "us", "nanosecond": "ns", "nanoseconds": "ns",
    "picosecond": "ps", "minute": "min", "hour": "h", "hertz": "Hz", "kilohertz": "kHz",
    "megahertz": "MHz", "gigahertz": "GHz",
}


class Unit:
    """scale * second**dim, with a printable symbol."""

    __slots__ = ("scale", "dim", "symbol")

    def __init__(self, scale, dim, symbol: str):
        self.scale = scale if isinstance(scale, Fraction) else Fraction(scale).limit_denominator(10**24)
        self.dim = Fraction(dim)
        self.symbol = symbol

    # -- algebra -----------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale * other.scale, self.dim + other.dim, _join(self.symbol, other.symbol, "*"))
        if isinstance(other, Quantity):
            return Quantity(other.m, self * other.u)
        if isinstance(other, Real):
            return Quantity(other, self)
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale / other.scale, self.dim - other.dim, _join(self.symbol, other.symbol, "/"))
        if isinstance(other, Quantity):
            return Quantity(1.0 / other.m, self / other.u)
        if isinstance(other, Real):
            return Quantity(1.0 / other, self)
        return NotImplemented

    def __rtruediv__(self, other):
        if isinstance(other, Real):
            return Quantity(other, Unit(1 / self.scale, -self.dim, _join("1", self.symbol, "/")))
        return NotImplemented

    def __pow__(self, p):
        p = Fraction(p).limit_denominator(64)
        if p.denominator == 1:
            scale = self.scale ** int(p)
        else:
            scale = Fraction(float(self.scale) ** float(p)).limit_denominator(10**24)
        return Unit(scale, self.dim * p, f"{self.symbol}**{p}" if p != 1 else self.symbol)

    def __eq__(self, other):
        if isinstance(other, Unit):
            return self.dim == other.dim and self.scale == other.scale
        return NotImplemented

    def __hash__(self):
        return hash((self.dim, self.scale))

    @property
    def dimensionless(self) -> bool:
        return self.dim == 0

    def __format__(self, spec):
        return self.symbol

    def __str__(self):
        return self.symbol

    def __repr__(self):
        return f"<Unit {self.symbol}>"


class Quantity:
    """magnitude * Unit"""

    __slots__ = ("m", "u")

    def __init__(self, magnitude, unit=None):
        if isinstance(magnitude, Quantity):
            unit = magnitude.u if unit is None else unit
            magnitude = magnitude.m
        if isinstance(unit, str):
            unit = ureg.unit(unit)
        if unit is None:
            unit = ureg.dimensionless
        if isinstance(magnitude, str):  # Quantity("16*ns")-style use is not needed; names only
            q = ureg(magnitude)
            magnitude, unit = q.m, q.u
        self.m = magnitude
        self.u = unit

    magnitude = property(lambda self: self.m)
    units = property(lambda self: self.u)

    def to(self, unit) -> "Quantity":
        if isinstance(unit, str):
            unit = ureg.unit(unit)
        if isinstance(unit, Quantity):
            unit = unit.u
        if unit.dim != self.u.dim:
            raise ValueError(f"cannot convert {self.u} to {unit}")
        return Quantity(_scale_mul(self.m, self.u.scale / unit.scale), unit)

    def _coerce(self, other):
        if isinstance(other, Quantity):
            return other
        if isinstance(other, Unit):
            return Quantity(1.0, other)
        if isinstance(other, Real):
            return Quantity(other, ureg.dimensionless)
        return None

    def __mul__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return Quantity(self.m * o.m, self.u * o.u)

    __rmul__ = __mul__

    def __truediv__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return Quantity(self.m / o.m, self.u / o.u)

    def __rtruediv__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return Quantity(o.m / self.m, o.u / self.u)

    def __floordiv__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        if o.u.dim != self.u.dim:
            raise ValueError("floor division needs compatible units")
        return Quantity(self.to(o.u).m // o.m, ureg.dimensionless)

    def __pow__(self, p):
        return Quantity(self.m ** float(p), self.u ** p)

    def _same(self, other):
        o = self._coerce(other)
        if o is None or o.u.dim != self.u.dim:
            raise ValueError(f"incompatible units {self.u} and {getattr(o, 'u', other)}")
        return o.to(self.u)

    def __add__(self, other):
        return Quantity(self.m + self._same(other).m, self.u)

    __radd__ = __add__

    def __sub__(self, other):
        return Quantity(self.m - self._same(other).m, self.u)

    def __rsub__(self, other):
        return Quantity(self._same(other).m - self.m, self.u)

    def __neg__(self):
        return Quantity(-self.m, self.u)

    def __float__(self):
        if self.u.dim != 0:
            raise TypeError(f"only dimensionless quantities convert to float, not {self.u}")
        return float(_scale_mul(self.m, self.u.scale))

    def __eq__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return o.u.dim == self.u.dim and _close(float(_scale_mul(self.m, self.u.scale)),
                                                float(_scale_mul(o.m, o.u.scale)))

    def __lt__(self, other):
        return self.m < self._same(other).m

    def __le__(self, other):
        return self.m <= self._same(other).m

    def __gt__(self, other):
        return self.m > self._same(other).m

    def __ge__(self, other):
        return self.m >= self._same(other).m

    def __hash__(self):
        return hash((self.u.dim, round(float(_scale_mul(self.m, self.u.scale)), 15)))

    def __str__(self):
        return f"{self.m} {self.u}"

    def __repr__(self):
        return f"<Quantity({self.m}, '{self.u}')>"


def _scale_mul(m, scale: Fraction):
    """m * scale with a single rounding (exact for integer-valued results)"""
    if scale == 1:
        return m
    if scale.denominator == 1:
        return m * scale.numerator
    if scale.numerator == 1:
        return m / scale.denominator
    return m * scale.numerator / scale.denominator


def _close(a: float, b: float) -> bool:
    return abs(a - b) <= 1e-12 * max(abs(a), abs(b), 1e-300)


def _join(a: str, b: str, op: str) -> str:
    if a in ("", "dimensionless"):
        return b if op == "*" else f"1/{b}"
    if b in ("", "dimensionless"):
        return a
    return f"{a}{op}{b}" if op == "*" else f"{a}/{b}"


class UnitRegistry:
    """``name in ureg`` / ``ureg(name)`` / ``ureg.Quantity`` like the pint registry the
    reference uses (only for the units a DSP chain needs)."""

    def __init__(self):
        self.dimensionless = Unit(Fraction(1), 0, "dimensionless")

    def __contains__(self, name) -> bool:
        return isinstance(name, str) and name in _UNITS

    def unit(self, name: str) -> Unit:
        if name not in _UNITS:
            raise KeyError(f"unknown unit {name!r}")
        scale, dim = _UNITS[name]
        return Unit(scale, dim, _SYMBOL.get(name, name))

    def __call__(self, name: str) -> Quantity:
        return Quantity(1.0, self.unit(name))

    def Quantity(self, value, unit=None) -> Quantity:  # noqa: N802 (pint spelling)
        if isinstance(value, str) and unit is None:
            return self(value)
        return Quantity(value, unit)

    def is_compatible_with(self, a, b) -> bool:
        ua, ub = as_unit(a), as_unit(b)
        return ua is not None and ub is not None and ua.dim == ub.dim


ureg = UnitRegistry()
unit_registry = ureg


def as_unit(x) -> Unit | None:
    """Unit of a Unit / Quantity / registry name; None for opaque labels."""
    x = from_foreign(x)
    if isinstance(x, Unit):
        return x
    if isinstance(x, Quantity):
        return x.u
    if isinstance(x, str) and x in ureg:
        return ureg.unit(x)
    return None


def is_in_registry(unit) -> bool:
    """Equivalent of the reference's ``is_in_pint`` (processing_chain.py:63-64)."""
    unit = from_foreign(unit)
    return isinstance(unit, (Unit, Quantity)) or bool(unit and unit in ureg)


def from_foreign(x):
    """Convert pint objects (if pint is installed and the caller used it) to ours."""
    cls = type(x)
    if cls.__module__.split(".")[0] == "pint":
        try:
            if hasattr(x, "magnitude"):
                base = x.to_base_units()
                dims = dict(base.units.dimensionality)
                dim = dims.pop("[time]", 0)
                if dims:
                    return str(x)
                return Quantity(float(base.magnitude), Unit(1.0, dim, "s" if dim == 1 else f"s**{dim}")).to(
                    Unit(1.0, dim, "s" if dim == 1 else f"s**{dim}")
                )
            q = 1 * x
            return from_foreign(q).u
        except Exception:
            return str(x)
    return x


def to_period_units(param: Quantity, period: Quantity) -> float:
    """Make ``param`` dimensionless by multiplying with the right power of the grid
    ``period`` (what the reference gets from ``pi_theorem``, processing_chain.py:1759-1766):
    a time becomes a number of samples, a frequency a number of cycles per sample, etc."""
    if period.u.dim == 0:
        raise ValueError("grid period is dimensionless")
    power = -param.u.dim / period.u.dim
    res = param * period ** power
    if res.u.dim != 0:
        raise ValueError(f"could not find valid conversion for {param}; grid period is {period}")
    return float(res)
