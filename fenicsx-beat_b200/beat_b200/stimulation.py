"""Stimulus (src/beat/stimulation.py:14-24) and define_stimulus (:210-272) without pint."""

from __future__ import annotations

from typing import NamedTuple

import numpy as np

from . import fem
from .units import CM_PER_UNIT, PerLength


class Stimulus(NamedTuple):
    expr: object  # fem.TimeWindow | fem.TimeFunction | fem.Separable | fem.Constant | float
    dZ: fem.Measure
    marker: int | None = None

    @property
    def dz(self):
        return self.dZ(self.marker)

    def assign(self, amp: float):
        self.expr.amplitude = amp


def compute_effective_dim(mesh: fem.Mesh, subdomain_data: fem.MeshTags) -> int:
    """Dimension used for the unit of the stimulus: subdomains of surfaces and lines are seen as slices of 3-D
    (stimulation.py:27-58)."""
    tdim = mesh.topology.dim
    if tdim not in (1, 2, 3):
        raise ValueError("Invalid mesh topology dimension")
    return subdomain_data.dim + (3 - tdim)


def get_dZ(mesh: fem.Mesh, subdomain_data: fem.MeshTags) -> fem.Measure:
    """``ds`` for facet tags, ``dx`` for cell tags (stimulation.py:61-108)."""
    tdim = mesh.topology.dim
    if subdomain_data.dim == tdim - 1:
        if tdim <= 1:
            raise ValueError("Invalid mesh topology dimension")
        return fem.Measure("ds", domain=mesh, subdomain_data=subdomain_data)
    if subdomain_data.dim == tdim:
        return fem.Measure("dx", domain=mesh, subdomain_data=subdomain_data)
    raise ValueError("Invalid subdomain data dimension")


def convert_chi(chi, mesh_unit: str) -> PerLength:
    """A plain number is taken in 1/mesh_unit (the reference's rule, stimulation.py:186-207); a PerLength keeps its unit."""
    if mesh_unit not in CM_PER_UNIT:
        raise ValueError(f"Invalid mesh unit {mesh_unit}")
    return chi if isinstance(chi, PerLength) else PerLength(float(chi), mesh_unit)


def amplitude_exponent(effective_dim: int) -> int:
    """A plain amplitude is in uA/cm^k with k = 1, 1, 2, 3 for effective dimension 0, 1, 2, 3 (stimulation.py:111-148)."""
    if effective_dim < 0 or effective_dim > 3:
        raise ValueError(f"Invalid effective dimension {effective_dim}. Must be 0, 1, 2 or 3.")
    return max(effective_dim, 1)


def define_stimulus(mesh: fem.Mesh, chi, time: fem.Constant, subdomain_data: fem.MeshTags, marker: int,
                    mesh_unit: str = "cm", duration: float = 2.0, amplitude: float = 500.0, start: float = 0.0) -> Stimulus:
    """amplitude [uA/cm^k] / chi expressed in uA/mesh_unit^(effective_dim-1), active for start <= time <= start + duration
    (stimulation.py:210-272 with the unit rules of :27-207 written out).

    ``chi``: a plain number is in 1/mesh_unit, as in the reference; ``conductivities.default_conductivities`` hands out
    ``PerLength(1400, "cm")`` like the reference's pint quantity, which is converted."""
    if mesh_unit not in CM_PER_UNIT:
        raise ValueError(f"Invalid mesh unit {mesh_unit}")
    effective_dim = compute_effective_dim(mesh, subdomain_data)
    k = amplitude_exponent(effective_dim)
    u = CM_PER_UNIT[mesh_unit]  # length of one mesh unit in cm
    chi_mesh = float(convert_chi(chi, mesh_unit).to(mesh_unit))  # 1/mesh_unit
    # A/chi = (amplitude/chi_mesh) uA mesh_unit / cm^k, and 1/cm^k = u^k / mesh_unit^k
    amp = amplitude / chi_mesh * u**k
    dZ = get_dZ(mesh, subdomain_data)
    expr = fem.TimeWindow(time, start, start + duration, amp)
    return Stimulus(dZ=dZ, marker=marker, expr=expr)


def generate_random_activation(mesh: fem.Mesh, time: fem.Constant, points: np.ndarray, delays: np.ndarray, stim_start: float = 0.0,
                               stim_duration: float = 2.0, stim_amplitude: float = 1.0, tol: float = 1e-12) -> fem.ExprSum:
    """Random spatio-temporal activation pattern (stimulation.py:279-363): the sum over the points of
    ``conditional(|x - p_i|_inf <= tol  and  start + delay_i <= time <= start + duration + delay_i, amplitude, 0)``.
    Returned as a sum of device-evaluated windows (one stimulus each); pass it as ``I_s`` or inside a ``Stimulus``."""
    points = np.atleast_2d(np.asarray(points, dtype=np.float64)) if len(points) else np.zeros((0, 3))
    assert len(points) == len(delays), "Points and delays must have the same length"
    gdim = mesh.geometry.x.shape[1] if mesh.geometry.x.ndim == 2 else 3

    def indicator(p):
        def g(x):  # near(X[k], p[k], tol) for every coordinate the point has (stimulation.py:275-276)
            ok = np.ones(np.asarray(x).shape[1], dtype=bool)
            for k in range(min(len(p), gdim, np.asarray(x).shape[0])):
                ok &= (x[k] >= p[k] - tol) & (x[k] <= p[k] + tol)
            return ok.astype(np.float64)

        return g

    return fem.ExprSum([fem.WindowedField(time, stim_start + float(d), stim_start + stim_duration + float(d), stim_amplitude, indicator(p))
                        for p, d in zip(points, delays)])
