"""The two unit-carrying values the hot path's set-up needs (the reference uses pint, src/beat/units.py; pint is not
installable here and nothing else of it is on the path): a length unit table and ``PerLength``, a float that remembers
the length unit of a surface-to-volume ratio such as ``1400 / cm``."""

from __future__ import annotations

CM_PER_UNIT = {"m": 100.0, "dm": 10.0, "cm": 1.0, "mm": 0.1, "um": 1.0e-4}  # length of one unit, in cm


class PerLength(float):
    """``value / unit`` (what ``1400.0 * ureg("cm**-1")`` is in the reference, conductivities.py:36).  Behaves as the plain
    number ``value`` in arithmetic; ``to(unit)`` converts."""

    unit: str

    def __new__(cls, value: float, unit: str = "cm"):
        if unit not in CM_PER_UNIT:
            raise ValueError(f"Invalid length unit {unit}")
        obj = super().__new__(cls, value)
        obj.unit = unit
        return obj

    def to(self, unit: str) -> "PerLength":
        if unit not in CM_PER_UNIT:
            raise ValueError(f"Invalid length unit {unit}")
        return PerLength(float(self) * CM_PER_UNIT[unit] / CM_PER_UNIT[self.unit], unit)

    def __repr__(self) -> str:
        return f"{float(self)!r} / {self.unit}"
