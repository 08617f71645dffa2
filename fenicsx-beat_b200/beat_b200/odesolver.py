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


class ODESystemSolver:
    """The reference's plain ODE driver (src/beat/odesolver.py:46-79): ``states[:] = fun(states, t0, parameters, dt)`` over
    an (num_states, num_points) array, here on the device with its own context (no mesh, no PDE).  ``states`` is
    held by reference like in the reference: it is uploaded before a step when the host copy was handed out since the
    last one, and read back lazily through the ``states`` property."""

    def __init__(self, fun: DeviceODE, states: np.ndarray, parameters: np.ndarray, missing_variables: np.ndarray | None = None,
                 monitor: BaseMonitor | None = None, v_index: int = 0, device: int | None = None):
        import os

        from ._lib import Context

        if not isinstance(fun, DeviceODE):
            raise TypeError("fun must be a device model handle (beat_b200.models.<model>.<scheme>); there is no CPU fallback")
        if missing_variables is not None:
            raise NotImplementedError("missing_variables (mechanics coupling) is outside the monodomain step")
        states = np.asarray(states, dtype=np.float64)
        if states.ndim != 2 or states.shape[0] != fun.num_states:
            raise ValueError(f"states must have shape ({fun.num_states}, num_points); got {states.shape}")
        self.fun, self.parameters, self.monitor = fun, parameters, monitor or NullMonitor()
        self.missing_variables = None
        self._ctx = Context(int(os.environ.get("MONO_DEVICE", "0")) if device is None else device)
        self._ctx.ode_create(fun.model_id, fun.scheme_id, states.shape[1], int(v_index), fun.num_states)
        self._mirror = _StateMirror(np.ascontiguousarray(states), self._ctx)
        self._params_raw: bytes | None = None

    @property
    def states(self) -> np.ndarray:
        return self._mirror.get()

    @property
    def num_points(self) -> int:
        return self._mirror.host.shape[1]

    @property
    def num_states(self) -> int:
        return self._mirror.host.shape[0]

    def _sync_parameters(self) -> None:
        p = np.asarray(self.parameters, dtype=np.float64)
        if p.ndim == 1:
            raw = p.tobytes()
            if raw != self._params_raw:
                self._ctx.ode_set_params(p, self.fun.derived(p))
                self._params_raw = raw
        else:
            if p.shape != (self.fun.num_parameters, self.num_points):
                raise ValueError(f"per-node parameters must have shape ({self.fun.num_parameters}, {self.num_points}); got {p.shape}")
            self._ctx.ode_set_params(p)  # may have been written by the caller: re-upload (pace_train.py:133-167)

    def step(self, t0: float, dt: float) -> None:  # odesolver.py:67-79
        with self.monitor.track_time("ode_total_step"):
            self._mirror.flush()
            self._sync_parameters()
            self._ctx.ode_step(t0, dt)
            self._mirror.mark_device_newer()

    def state_row(self, index: int) -> np.ndarray:
        """One state (e.g. the membrane potential) of all points, without moving the other rows."""
        self._mirror.flush()
        return self._ctx.ode_get_state_row(int(index))

    def set_state_row(self, index: int, values: np.ndarray) -> None:
        self._mirror.flush()
        self._ctx.ode_set_state_row(int(index), np.ascontiguousarray(values, dtype=np.float64))
        self._mirror.mark_device_newer()


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
        self._n = int(self.v_ode.x.array_ro.size)  # owned + ghosts, odesolver.py:189-190
        full = (self.num_states, self._n)
        if np.shape(self.init_states) == full:
            values = np.array(self.init_states, dtype=np.float64, order="C")
        else:
            values = np.zeros(full)
            values.T[:] = self.init_states
        self._ctx = ctx = v_pde.function_space.mesh.device_context()
        ctx.ode_create(fun.model_id, fun.scheme_id, self._n, self.v_index, num_states)
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
            if p.shape[1] != self._n:
                raise ValueError(f"per-node parameters must have shape (num_parameters, {self._n}); got {p.shape}")
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


class DolfinMultiODESolver(DolfinODESolver):
    """Per-region cell models (src/beat/odesolver.py:228-354): ``markers`` is a P1 function holding an integer
    region id per dof; ``init_states`` / ``parameters`` / ``fun`` / ``num_states`` / ``v_index`` are dicts keyed by it.

    On the device this is ONE kernel launch over all nodes: a per-node region index selects one of a few parameter
    sets (shared parameters + host-evaluated derived constants) kept in a small table (the reference loops over the
    regions and runs one NumPy update per region).  That needs the same cell model in every region - the case of the
    reference's demos (ToR-ORd / TP06 with endo / mid / epi parameter sets); different models per region raise
    NotImplementedError."""

    def __new__(cls, v_ode=None, v_pde=None, markers=None, init_states=None, parameters=None, fun=None, num_states=None,
                v_index=None, monitor=None, **kw):
        # different cell MODELS (or schemes / v_index) per region: one device ODE stage per region (MixedModelODESolver)
        if cls is DolfinMultiODESolver and isinstance(fun, dict) and isinstance(v_index, dict) and init_states is not None:
            funs = [fun[m] for m in init_states]
            if all(isinstance(f, DeviceODE) for f in funs) and (
                    len({(f.model_id, f.scheme_id) for f in funs}) != 1 or len({v_index[m] for m in init_states}) != 1):
                return MixedModelODESolver(v_ode=v_ode, v_pde=v_pde, markers=markers, init_states=init_states, parameters=parameters,
                                           fun=fun, num_states=num_states, v_index=v_index, monitor=monitor, **kw)
        return super().__new__(cls)

    def __init__(self, v_ode: fem.Function, v_pde: fem.Function, markers: fem.Function, init_states: dict, parameters: dict,
                 fun: dict, num_states: dict, v_index: dict, monitor: BaseMonitor | None = None):
        if v_ode.x.array_ro.size != markers.x.array_ro.size:
            raise RuntimeError("Marker and voltage need to be in the same function space")  # odesolver.py:241-242
        self._marker_values = tuple(init_states.keys())
        funs = [fun[m] for m in self._marker_values]
        if any(not isinstance(f, DeviceODE) for f in funs):
            raise TypeError("fun[marker] must be device model handles (there is no CPU fallback)")
        if len({(f.model_id, f.scheme_id) for f in funs}) != 1 or len({v_index[m] for m in self._marker_values}) != 1:
            raise NotImplementedError("one cell model over all regions here; DolfinMultiODESolver(...) dispatches mixed models "
                                      "to MixedModelODESolver")
        marr = np.asarray(markers.x.array_ro)
        self.markers = markers
        self._inds = {m: marr == m for m in self._marker_values}
        self._num_points_m = {m: int(w.sum()) for m, w in self._inds.items()}
        covered = np.zeros(marr.size, dtype=bool)
        for w in self._inds.values():
            covered |= w
        if not covered.all():
            raise ValueError("every dof needs a marker that has a cell model")
        f0 = funs[0]
        ns = f0.num_states
        n = marr.size
        values = np.zeros((ns, n))
        self._region_of_node = np.zeros(n, dtype=np.int32)
        for k, m in enumerate(self._marker_values):
            if num_states[m] != ns:
                raise ValueError(f"num_states[{m}] = {num_states[m]} but the model has {ns} states")
            init = np.asarray(init_states[m], dtype=np.float64)
            values[:, self._inds[m]] = init if init.shape == (ns, self._num_points_m[m]) else init.reshape(ns, 1)
            self._region_of_node[self._inds[m]] = k
            if np.ndim(parameters[m]) != 1:
                raise NotImplementedError("per-region parameters must be 1-D vectors (one set per region)")
        self._region_parameters = parameters
        self._region_raw: dict | None = None
        super().__init__(v_ode=v_ode, v_pde=v_pde, init_states=values, parameters=parameters, fun=f0, num_states=ns,
                         v_index=v_index[self._marker_values[0]], monitor=monitor)

    def _sync_parameters(self) -> None:
        # the per-region vectors are held by reference and may be mutated in place between steps (pace_train.py:224)
        raw = {m: np.asarray(self._region_parameters[m], dtype=np.float64).tobytes() for m in self._marker_values}
        if raw != self._region_raw:
            p = np.stack([np.asarray(self._region_parameters[m], dtype=np.float64) for m in self._marker_values])
            d = np.stack([self.fun.derived(row) for row in p])
            self._ctx.ode_set_region_params(p, d, self._region_of_node if self._region_raw is None else None)
            self._region_raw = raw
        self._params_dirty = False

    # reference surface with a marker argument (odesolver.py:292-303)
    def values(self, marker: int) -> np.ndarray:  # type: ignore[override]
        return self._mirror.get()[:, self._inds[marker]]

    def num_parameters(self, marker: int) -> int:  # type: ignore[override]
        return len(self._region_parameters[marker])

    def shape(self, marker: int):  # type: ignore[override]
        return (self.num_states, self._num_points_m[marker])

    def num_points(self, marker: int) -> int:  # type: ignore[override]
        return self._num_points_m[marker]

    @property
    def full_values(self) -> np.ndarray:
        return self._mirror.get()


class MixedModelODESolver(BaseDolfinODESolver):
    """``DolfinMultiODESolver`` with DIFFERENT cell models per region (src/beat/odesolver.py:228-354 allows any callable per
    marker; demos/lv_endocardial.py:203-257 uses one model with three parameter sets, which the single-launch
    DolfinMultiODESolver above serves).  One device ODE stage (``ODESystemSolver``, its own context and compact state array)
    per region; the membrane potential is exchanged with the PDE stage through the host mirrors of ``v_ode`` / ``v_pde`` -
    one row per region and direction per hand-off, so this path costs PCIe traffic every step and runs through the
    splitting solver's protocol sequence (monodomain_solver.py:66-113), not the fused device step.  It exists so that the
    reference's interface is complete; it is a composition of device primitives that are each covered by the GPU tests and
    is itself exercised on the CPU with a stand-in for the device stage (tests/test_host_logic.py)."""

    def __init__(self, v_ode: fem.Function, v_pde: fem.Function, markers: fem.Function, init_states: dict, parameters: dict,
                 fun: dict, num_states: dict, v_index: dict, monitor: BaseMonitor | None = None, system_solver=None):
        if v_ode.x.array_ro.size != markers.x.array_ro.size:
            raise RuntimeError("Marker and voltage need to be in the same function space")  # odesolver.py:241-242
        if v_ode.x.array_ro.size != v_pde.x.array_ro.size:
            raise NotImplementedError("v_ode and v_pde must be the same P1 space (identity projection, utils.py:52-54)")
        make = system_solver or ODESystemSolver
        self.v_ode, self.v_pde, self.markers = v_ode, v_pde, markers
        self.init_states, self.parameters, self.fun = init_states, parameters, fun
        self.num_states, self.v_index = dict(num_states), {m: int(v_index[m]) for m in init_states}
        self.monitor = monitor or NullMonitor()
        self._marker_values = tuple(init_states.keys())
        marr = np.asarray(markers.x.array_ro)
        self._inds = {m: np.nonzero(marr == m)[0] for m in self._marker_values}
        if sum(len(w) for w in self._inds.values()) != marr.size:
            raise ValueError("every dof needs a marker that has a cell model")
        self._odes = {}
        for m in self._marker_values:
            shape = (int(num_states[m]), len(self._inds[m]))
            init = np.asarray(init_states[m], dtype=np.float64)
            values = np.array(init, order="C") if init.shape == shape else np.repeat(init.reshape(shape[0], 1), shape[1], axis=1)
            self._odes[m] = make(fun=fun[m], states=values, parameters=parameters[m], monitor=self.monitor, v_index=self.v_index[m])
        self._pde = getattr(v_pde, "_owner", None)

    # ---- the five hand-offs of the ODESolver protocol ---------------------------------------------------------------
    def step(self, t0: float, dt: float) -> None:
        with self.monitor.track_time("total_ode_step"):
            for m, ode in self._odes.items():
                with self.monitor.track_time(f"marker_{m}_ode_step"):
                    ode.step(t0=t0, dt=dt)

    def to_dolfin(self) -> None:  # states[v_index] -> v_ode, region by region
        arr = self.v_ode.x.array
        for m, ode in self._odes.items():
            arr[self._inds[m]] = ode.state_row(self.v_index[m])

    def from_dolfin(self) -> None:  # v_ode -> states[v_index]
        arr = self.v_ode.x.array_ro
        for m, ode in self._odes.items():
            ode.set_state_row(self.v_index[m], arr[self._inds[m]])

    def ode_to_pde(self) -> None:  # identity projection between the two P1 functions, on the host mirrors
        self.v_pde.x.array[:] = self.v_ode.x.array_ro

    def pde_to_ode(self) -> None:
        self.v_ode.x.array[:] = self.v_pde.x.array_ro

    # ---- reference surface with a marker argument (odesolver.py:292-303) ----------------------------------------------
    def values(self, marker: int) -> np.ndarray:
        return self._odes[marker].states

    def num_parameters(self, marker: int) -> int:
        return len(self.parameters[marker])

    def shape(self, marker: int) -> tuple[int, int]:
        return (self.num_states[marker], len(self._inds[marker]))

    def num_points(self, marker: int) -> int:
        return len(self._inds[marker])

    @property
    def full_values(self) -> np.ndarray:
        sizes = set(self.num_states.values())
        if len(sizes) != 1:
            raise RuntimeError(f"Cannot get full values since states are not of equal size. Have {self.num_states=}, "
                               "use .values(marker) instead")
        out = np.zeros((sizes.pop(), self.markers.x.array_ro.size))
        for m, ode in self._odes.items():
            out[:, self._inds[m]] = ode.states
        return out

    def assign_all_states(self, functions: list[fem.Function]) -> None:
        n = self.num_states[self._marker_values[0]]
        assert len(functions) == n, "Number of functions must match number of states"
        for index, f in enumerate(functions):
            for m, ode in self._odes.items():
                f.x.array[self._inds[m]] = ode.states[index, :]

    def states_to_dolfin(self, names: list[str] | None = None) -> list[fem.Function]:
        n = self.num_states[self._marker_values[0]]
        if names is not None:
            assert len(names) == n, f"Number of names must match number of states, got {len(names)} names, but number of states is {n}"
        else:
            names = [f"state_{i}" for i in range(n)]
        functions = [fem.Function(self.v_ode.function_space, name=name) for name in names]
        self.assign_all_states(functions)
        return functions
