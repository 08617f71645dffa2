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

    # fp64-pipe instructions per operation of the generated code (csrc/ode_math.cuh "fast" binding):
    # divide = reciprocal seed + 4 FMA + mul + 2 FMA; exp/log = libdevice polynomial kernels.
    FP64_WEIGHTS = {"add": 1, "mul": 1, "div": 7, "exp": 17, "log": 25, "sqrt": 12, "pow": 60, "floor": 1, "cmp": 1}

    def fp64_instr_per_node(self) -> float:
        """Estimated fp64-pipe instructions one node-step issues (roofline numerator of K1)."""
        return float(sum(self.FP64_WEIGHTS.get(k, 0) * v for k, v in self.op_counts.items()))

    def __call__(self, *args, **kwargs):
        raise RuntimeError(
            f"{self.model_tag}.{self.scheme} is a device kernel handle and cannot be evaluated on the CPU "
            "(no CPU fallback): pass it as `fun=` to beat_b200.odesolver.DolfinODESolver"
        )
