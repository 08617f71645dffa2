"""ODE stage on the device: the reference's DolfinODESolver (src/beat/odesolver.py:135-225) with the same
constructor, attributes and methods.  ``fun`` must be a :class:`DeviceODE` handle from
``beat_b200.models`` (a GPU cannot run an arbitrary Python callable and there is no CPU fallback).

Ownership follows the reference: ``init_states`` is copied (odesolver.py:149-153); ``parameters`` is held
by reference and may be mutated by the user between steps (demos/pace_train.py:224) - shared (1-D)
parameters are compared against the last uploaded copy every step, per-node (2-D) parameters are
re-uploaded after any access through ``ode.parameters`` (or ``mark_parameters_dirty()``).
"""

from __future__ import annotations

import abc
import logging
from typing import Any

import numpy as np

from . import fem
from .device_model import DeviceODE
from .telemetry import BaseMonitor, NullMonitor

logger = logging.getLogger(__name__)


class BaseDolfinODESolver(abc.ABC):
    v_ode: fem.Function
    v_pde: fem.Function
    _metadata: dict[str, Any] | None = None

    @abc.abstractmethod
    def to_dolfin(self) -> None: ...

    @abc.abstractmethod
    def from_dolfin(self) -> None: ...

    @abc.abstractmethod
    def ode_to_pde(self) -> None: ...

    @abc.abstractmethod
    def pde_to_ode(self) -> None: ...

    @abc.abstractmethod
    def step(self, t0: float, dt: float) -> None: ...


class _StateMirror:
    """Host mirror of the (num_states, num_points) state array with the same coherence rules as
    fem.Vector: reading ``values`` downloads if the device is newer and schedules a re-upload."""

    def __init__(self, host: np.ndarray, ctx):
        self.host, self.ctx = host, ctx
        self.device_newer = False
        self.host_dirty = True

    def get(self) -> np.ndarray:
        if self.device_newer:
            self.ctx.ode_get_states(self.host)
            self.device_newer = False
        self.host_dirty = True
        return self.host

    def flush(self):
        if self.host_dirty:
            self.ctx.ode_set_states(self.host)
            self.host_dirty = False

    def mark_device_newer(self):
        self.device_newer = True
        self.host_dirty = False


class DolfinODESolver(BaseDolfinODESolver):
    def __init__(self, v_ode: fem.Function, v_pde: fem.Function, init_states: np.ndarray, parameters: np.ndarray | None,
                 fun: DeviceODE, num_states: int, v_index: int = 0, missing_variables: np.ndarray | None = None,
                 num_missing_variables: int = 0, monitor: BaseMonitor | None = None):
        if not isinstance(fun, DeviceODE):
            raise TypeError(
                "fun must be a device model handle (e.g. beat_b200.models.tp06.generalized_rush_larsen); "
                "arbitrary Python callables cannot run on the GPU and there is no CPU fallback"
            )
        if missing_variables is not None or num_missing_variables:
            raise NotImplementedError("missing_variables (mechanics coupling) is outside the monodomain step")
        if v_ode.x.array_ro.size != v_pde.x.array_ro.size:
            raise NotImplementedError("v_ode and v_pde must be the same P1 space (identity projection, utils.py:52-54)")
        if num_states != fun.num_states:
            raise ValueError(f"model {fun.model_tag} has {fun.num_states} states, got num_states={num_states}")
        self.v_ode, self.v_pde = v_ode, v_pde
        self.init_states, self._parameters, self.fun = init_states, parameters, fun
        self.num_states, self.v_index = num_states, int(v_index)
        self.missing_variables, self.num_missing_variables = None, 0
        self.monitor = monitor or NullMonitor()

        # odesolver.py:148-153
        if np.shape(self.init_states) == self.shape:
            values = np.array(self.init_states, dtype=np.float64, order="C")
        else:
            values = np.zeros(self.shape)
            values.T[:] = self.init_states
        self._ctx = ctx = v_pde.function_space.mesh.device_context()
        ctx.ode_create(fun.model_id, fun.scheme_id, self.num_points, self.v_index, num_states)
        self._mirror = _StateMirror(values, ctx)
        self._mirror.flush()
        self._params_uploaded: bytes | None = None
        self._params_dirty = True
        self._sync_parameters()
        self.v_ode.x.bind(ctx.get_v_ode, ctx.set_v_ode, push_now=False, sync=ctx.sync)
        self._pde = getattr(v_pde, "_owner", None)

    # ---- parameters ------------------------------------------------------------------------------
    @property
    def parameters(self):
        if self._parameters is not None and np.ndim(self._parameters) == 2:
            self._params_dirty = True  # the caller may write into the per-node table
        return self._parameters

    @parameters.setter
    def parameters(self, value):
        self._parameters = value
        self._params_dirty = True

    def mark_parameters_dirty(self) -> None:
        self._params_dirty = True

    def _sync_parameters(self) -> None:
        p = self._parameters
        if p is None:
            raise ValueError("parameters=None: the compiled cell models need their parameter vector")
        p = np.asarray(p, dtype=np.float64)
        if p.ndim == 1:
            raw = p.tobytes()  # users mutate the shared vector in place (pace_train.py:224): compare, cheaply
            if self._params_dirty or raw != self._params_uploaded:
                self._ctx.ode_set_params(p, self.fun.derived(p))
                self._params_uploaded = raw
        elif self._params_dirty:
            if p.shape[1] != self.num_points:
                raise ValueError(f"per-node parameters must have shape (num_parameters, {self.num_points}); got {p.shape}")
            self._ctx.ode_set_params(p)
        self._params_dirty = False

    # ---- reference surface -----------------------------------------------------------------------
    @property
    def values(self) -> np.ndarray:
        return self._mirror.get()

    @property
    def full_values(self) -> np.ndarray:
        return self._mirror.get()

    @property
    def num_parameters(self) -> int:
        return len(self._parameters)

    @property
    def shape(self) -> tuple[int, int]:
        return (self.num_states, self.num_points)

    @property
    def num_points(self) -> int:
        return self.v_ode.x.array_ro.size  # owned + ghosts, odesolver.py:189-190

    def _flush_host(self) -> None:
        self._mirror.flush()
        self.v_ode.x.flush_to_device()
        self.v_pde.x.flush_to_device()
        self._sync_parameters()

    def step(self, t0: float, dt: float) -> None:  # odesolver.py:67-79,192-193
        with self.monitor.track_time("ode_total_step"):
            self._flush_host()
            self._ctx.ode_step(t0, dt)
            self._mirror.mark_device_newer()

    def to_dolfin(self) -> None:  # odesolver.py:164-166
        self._flush_host()
        self._ctx.ode_to_dolfin()
        self.v_ode.x.mark_device_newer()

    def from_dolfin(self) -> None:  # odesolver.py:168-170
        self._flush_host()
        self._ctx.ode_from_dolfin()
        self._mirror.mark_device_newer()

    def ode_to_pde(self) -> None:  # odesolver.py:101-107
        self._flush_host()
        self._ctx.ode_to_pde()
        self.v_pde.x.mark_device_newer()

    def pde_to_ode(self) -> None:  # odesolver.py:109-115
        self._flush_host()
        self._ctx.pde_to_ode()
        self.v_ode.x.mark_device_newer()

    def assign_all_states(self, functions: list[fem.Function]) -> None:  # odesolver.py:195-199
        vals = self._mirror.get()
        assert len(functions) == vals.shape[0], "Number of functions must match number of states"
        for index, f in enumerate(functions):
            f.x.array[:] = vals[index, :]

    def states_to_dolfin(self, names: list[str] | None = None) -> list[fem.Function]:  # odesolver.py:201-225
        V = self.v_ode.function_space
        num_states = self.num_states
        if names is not None:
            msg = f"Number of names must match number of states, got {len(names)} names, but number of states is {num_states}"
            assert len(names) == num_states, msg
        else:
            names = [f"state_{i}" for i in range(num_states)]
        functions = [fem.Function(V, name=name) for name in names]
        self.assign_all_states(functions)
        return functions
