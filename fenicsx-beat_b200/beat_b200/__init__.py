"""beat_b200: B200-native drop-in for fenicsx-beat's operator-split monodomain step.

    import beat_b200 as beat

mirrors the names the reference exports for this path (src/beat/__init__.py:1-63): MonodomainModel,
MonodomainSplittingSolver, odesolver.DolfinODESolver, Stimulus, the monitors, conductivities, geometry,
stimulation.  ``beat_b200.fem`` stands in for the few dolfinx/ufl objects those signatures take, and
``beat_b200.models`` holds the compiled cell models (device handles with a gotranx-module-like surface).
"""

from . import conductivities, fem, geometry, models, monodomain_model, monodomain_solver, odesolver, single_cell, stimulation, telemetry, units
from .monodomain_model import MonodomainModel
from .monodomain_solver import MonodomainSplittingSolver
from .stimulation import Stimulus
from .telemetry import BaseMonitor, NullMonitor, PerformanceMonitor

base_model = monodomain_model  # the reference keeps Stimulus/Status/Results reachable via beat.base_model
monodomain_model.Stimulus = Stimulus

__all__ = [
    "MonodomainModel", "MonodomainSplittingSolver", "Stimulus", "BaseMonitor", "NullMonitor", "PerformanceMonitor",
    "odesolver", "single_cell", "conductivities", "geometry", "stimulation", "telemetry", "units", "fem", "models", "base_model",
]
__version__ = "0.1.0"
