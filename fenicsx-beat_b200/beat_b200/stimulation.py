"""Stimulus (src/beat/stimulation.py:14-24) and define_stimulus (:210-272) without pint."""

from __future__ import annotations

from typing import NamedTuple

from . import fem


class Stimulus(NamedTuple):
    expr: object  # fem.TimeWindow | fem.TimeFunction | fem.Separable | fem.Constant | float
    dZ: fem.Measure
    marker: int | None = None

    @property
    def dz(self):
        return self.dZ(self.marker)

    def assign(self, amp: float):
        self.expr.amplitude = amp


_TO_MESH_UNIT = {"m": 100.0, "dm": 10.0, "cm": 1.0, "mm": 0.1}  # 1 cm expressed in the mesh unit, inverted below


def define_stimulus(mesh: fem.Mesh, chi: float, time: fem.Constant, subdomain_data: fem.MeshTags, marker: int,
                    mesh_unit: str = "cm", duration: float = 2.0, amplitude: float = 500.0, start: float = 0.0) -> Stimulus:
    """amplitude [uA/cm^effective_dim] / chi [1/cm] expressed in uA/mesh_unit^(effective_dim-1), active for
    start <= time <= start + duration (stimulation.py:264-272; unit rules :27-207).

    effective dimension: subdomain dimension seen as a slice of 3-D (stimulation.py:27-58)."""
    if mesh_unit not in _TO_MESH_UNIT:
        raise ValueError(f"Invalid mesh unit {mesh_unit}")
    tdim = mesh.topology.dim
    effective_dim = subdomain_data.dim + (3 - tdim)
    if effective_dim < 0 or effective_dim > 3:
        raise ValueError("Invalid effective dimension")
    cm_per_unit = _TO_MESH_UNIT[mesh_unit]  # length of one mesh unit in cm
    # A/chi has unit uA/cm^(effective_dim-1); 1/cm^(k) = cm_per_unit^k / mesh_unit^k
    amp = amplitude / chi * cm_per_unit ** (effective_dim - 1)
    kind = "dx" if subdomain_data.dim == tdim else "ds"
    dZ = fem.Measure(kind, domain=mesh, subdomain_data=subdomain_data)
    expr = fem.TimeWindow(time, start, start + duration, amp)
    return Stimulus(dZ=dZ, marker=marker, expr=expr)
