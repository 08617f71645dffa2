"""Conductivity sets and tensors (src/beat/conductivities.py:29-118), without pint: values are plain
floats in the units stated in the names; the unit conversions the reference does with pint are written
out."""

from __future__ import annotations

from typing import NamedTuple

import numpy as np

from .fem import Constant
from .units import PerLength


def default_conductivities(name: str = "Niederer") -> dict[str, float]:
    """g_* in S/m (plain floats), chi as ``PerLength(.., "cm")`` (the reference returns pint quantities in these units)."""
    if name == "Niederer":
        return {"g_il": 0.17, "g_it": 0.019, "g_el": 0.62, "g_et": 0.24, "chi": PerLength(1400.0, "cm")}
    if name == "Bishop":
        return {"g_il": 0.34, "g_it": 0.060, "g_el": 0.12, "g_et": 0.08, "chi": PerLength(1400.0, "cm")}
    if name == "Potse":  # mS/cm -> S/m is a factor 0.1
        return {"g_il": 0.3, "g_it": 0.03, "g_el": 0.3, "g_et": 0.12, "chi": PerLength(800.0, "cm")}
    raise ValueError(f"Unknown conductivity tensor {name}")


class Conductivities(NamedTuple):
    s_l: float
    s_t: float


def get_harmonic_mean_conductivity(chi: float, g_il: float = 0.17, g_it: float = 0.019, g_el: float = 0.62,
                                   g_et: float = 0.24) -> Conductivities:
    """sigma = g_i g_e / (g_i + g_e) [S/m], scaled by 1/chi and expressed in uA/mV (conductivities.py:63-98):
    S/m / (1/cm) = S/m * 0.01 m = 0.01 S = 10 uA/mV.  ``chi``: PerLength, or a plain number taken in 1/cm (the reference
    cannot convert a unit-less chi here at all: pint raises)."""
    chi = float(chi.to("cm")) if isinstance(chi, PerLength) else float(chi)

    def harmonic_mean(a, b):
        return a * b / (a + b)

    sigma_l = harmonic_mean(g_il, g_el)
    sigma_t = harmonic_mean(g_it, g_et)
    return Conductivities(sigma_l / chi * 10.0, sigma_t / chi * 10.0)


def conductivity_tensor(s_l: float, s_t: float, f0) -> np.ndarray:
    """M = s_l f0 (x) f0 + s_t (I - f0 (x) f0) (conductivities.py:101-104).  f0: Constant / array (dim,)
    for a constant fibre, or (ncell, dim) for a cell-wise fibre field -> (ncell, dim, dim)."""
    f = np.asarray(f0.value if isinstance(f0, Constant) else f0, dtype=np.float64)
    dim = f.shape[-1]
    outer = f[..., :, None] * f[..., None, :]
    return s_l * outer + s_t * (np.eye(dim) - outer)


def define_conductivity_tensor(chi: float, f0, g_il: float = 0.17, g_it: float = 0.019, g_el: float = 0.62,
                               g_et: float = 0.24) -> np.ndarray:
    if f0 is None:
        raise ValueError("f0 must be provided")
    s_l, s_t = get_harmonic_mean_conductivity(chi, g_il, g_it, g_el, g_et)
    return conductivity_tensor(s_l, s_t, f0)
