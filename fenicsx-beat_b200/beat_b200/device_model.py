"""Handle that names a compiled cell-model kernel.

The reference passes an arbitrary Python callable ``fun(states, t, parameters, dt)`` to its ODE solver
(src/beat/odesolver.py:46-79).  A GPU cannot run a Python callable, and there is no CPU fallback, so the
drop-in takes one of these handles instead: the generated model modules expose them under the same
names a gotranx-generated module uses (``generalized_rush_larsen``, ``forward_explicit_euler``).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import numpy as np


@dataclass(frozen=True)
class DeviceODE:
    model_id: int
    model_tag: str
    scheme_id: int
    scheme: str
    num_states: int
    num_parameters: int
    derived: Callable[[np.ndarray], np.ndarray] = field(repr=False, compare=False)
    op_counts: dict = field(default_factory=dict, repr=False, compare=False)

    def __call__(self, *args, **kwargs):
        raise RuntimeError(
            f"{self.model_tag}.{self.scheme} is a device kernel handle and cannot be evaluated on the CPU "
            "(no CPU fallback): pass it as `fun=` to beat_b200.odesolver.DolfinODESolver"
        )
