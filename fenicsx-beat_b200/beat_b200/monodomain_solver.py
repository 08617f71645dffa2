"""MonodomainSplittingSolver (src/beat/monodomain_solver.py:26-116) on the device.

``step((t0, t1))`` keeps the reference's semantics - Godunov (theta = 1) or Strang (theta != 1) splitting,
ODE time arguments t0 and t0 + theta*dt, PDE stimulus at t0 + theta_pde*dt - but runs as ONE fused
device call (mono_split_step: cell-model kernel -> persistent diffusion kernel [-> cell-model kernel]),
with the v_ode / v_pde / v_ hand-off copies of :72-97 fused away.  Any ODE backend that is not a device
DolfinODESolver falls back to the reference's literal call sequence over the ODESolver protocol (:14-23).
"""

from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Protocol

import numpy as np

from .monodomain_model import MonodomainModel
from .odesolver import DolfinODESolver
from .telemetry import BaseMonitor, NullMonitor, PerformanceMonitor

logger = logging.getLogger(__name__)
EPS = 1e-12


class ODESolver(Protocol):
    """What the splitting solver needs from an ODE backend (monodomain_solver.py:14-23): advance the cell states, and
    move the membrane potential between the state array, its function in the ODE space and the PDE space."""

    def step(self, t0: float, dt: float) -> None: ...
    def to_dolfin(self) -> None: ...          # states[v_index] -> v_ode
    def from_dolfin(self) -> None: ...        # v_ode -> states[v_index]
    def ode_to_pde(self) -> None: ...         # v_ode -> v_pde
    def pde_to_ode(self) -> None: ...         # v_pde -> v_ode


@dataclass
class MonodomainSplittingSolver:
    pde: MonodomainModel
    ode: ODESolver
    theta: float = 1.0
    monitor: BaseMonitor = field(default_factory=NullMonitor)

    def __post_init__(self) -> None:  # monodomain_solver.py:33-37
        self.ode.to_dolfin()
        self.ode.ode_to_pde()
        self.pde.assign_previous()
        self._fused = isinstance(self.ode, DolfinODESolver) and getattr(self.ode, "_ctx", None) is self.pde._ctx
        self._timed = isinstance(self.monitor, PerformanceMonitor)
        if self._fused and self._timed:
            self.pde._ctx.stage_timing(True)

    def solve(self, interval, dt=None):  # monodomain_solver.py:39-51
        """Fixed-size steps over ``interval``; a step is taken while its END does not pass T (+1e-12), and every end time
        is previous end + dt (the reference's accumulation, so the same number of steps comes out)."""
        begin, end = interval
        width = (end - begin) if dt is None else dt
        left = begin
        while (right := left + width) < end + EPS:
            self.step((left, right))
            left = right

    def _prepare_fused(self, t0: float, t1: float) -> None:
        pde, ode = self.pde, self.ode
        dt = t1 - t0
        pde.time.value = t0 + pde.parameters["theta"] * dt  # base_model.py:216-223
        if not abs(dt - float(pde._timestep)) < 1.0e-12:
            pde._timestep.value = dt
        pde._push_stimulus_amplitudes()
        ode._flush_host()
        pde._flush_host()

    def _mark_fused(self) -> None:
        pde, ode = self.pde, self.ode
        ode._mirror.mark_device_newer()
        pde.state.x.mark_device_newer()
        # post-condition of the reference's step (monodomain_solver.py:88-97): v_ode == v == v_ ; whoever is read
        # second on the host copies from the first instead of crossing PCIe again
        ode.v_ode.x.mark_device_newer(twin_of=pde.state.x)
        pde.v_.x.mark_device_newer(twin_of=pde.state.x)

    def step(self, interval):  # monodomain_solver.py:53-116
        theta = self.theta
        t0, t1 = interval
        if not self._fused:
            return self._step_protocol(interval)
        with self.monitor.track_time("total_step"):
            self._prepare_fused(t0, t1)
            self.pde._ctx.split_step(t0, t1, theta)
            self._mark_fused()
        if self._timed:
            self.pde.monitor.record_ksp(self.pde.ksp) if self.pde.monitor is not self.monitor else None
            self.monitor.record_ksp(self.pde.ksp)
            st = self.pde._ctx.stage_times_ms(reset=True)
            self.monitor.record_device_stages({"ode_step": st["ode_ms"] * 1e-3, "pde_step": st["pde_ms"] * 1e-3})
        self.monitor.advance_step(t0, t1)

    def solve_on_device(self, t0: float, dt: float, nsteps: int) -> None:
        """``nsteps`` fixed-dt steps without returning to Python between them (mono_split_solve): the
        reference's ``solve`` loop minus the per-step interpreter cost.  Only stimuli whose amplitude is constant over the
        call can stay on the device for it (``TimeWindow``: the window test is part of the RHS kernel; plain constants).
        A source whose amplitude the HOST evaluates every step (``fem.TimeFunction`` / ``fem.Separable`` h(t)) would be
        frozen at its first value, so with any of those the call takes the per-step path - same result as ``solve``."""
        if not self._fused:
            raise RuntimeError("solve_on_device needs a device DolfinODESolver sharing the PDE's context")
        if nsteps <= 0:
            return
        if self.pde.has_host_evaluated_sources():
            t = t0
            for _ in range(nsteps):
                self.step((t, t + dt))
                t = t + dt
            return
        self._prepare_fused(t0, t0 + dt)
        self.pde._ctx.split_solve(t0, dt, nsteps, self.theta)
        self.pde.time.value = t0 + (nsteps - 1) * dt + self.pde.parameters["theta"] * dt
        self._mark_fused()
        t = t0
        for _ in range(nsteps):  # the monitors count steps as the reference's loop does (telemetry.py:86-92)
            self.monitor.advance_step(t, t + dt)
            t = t + dt

    def _step_protocol(self, interval):
        """Any other ODE backend (the 5-method ODESolver protocol): the hand-offs of the reference's step, one timed label
        each (monodomain_solver.py:66-113), as data - first half, then either the Godunov or the Strang tail."""
        t0, t1 = interval
        dt = t1 - t0
        ode, pde, theta = self.ode, self.pde, self.theta
        godunov = bool(np.isclose(theta, 1.0))
        first_half = (
            ("ode_step", lambda: ode.step(t0=t0, dt=theta * dt)),
            ("ode_to_dolfin", ode.to_dolfin),
            ("ode_to_pde", ode.ode_to_pde),
            ("pde_assign_previous_before", pde.assign_previous),
            ("pde_step", lambda: pde.step((t0, t1))),
            ("pde_to_ode", ode.pde_to_ode),
            ("ode_from_dolfin", ode.from_dolfin),
        )
        tail = (("pde_assign_previous_after", pde.assign_previous),) if godunov else (
            ("corrective_ode_step", lambda: ode.step(t0 + theta * dt, (1.0 - theta) * dt)),
            ("corrective_ode_to_dolfin", ode.to_dolfin),
            ("corrective_ode_to_pde", ode.ode_to_pde),
            ("corrective_pde_assign_previous", pde.assign_previous),
        )
        with self.monitor.track_time("total_step"):
            for label, action in first_half + tail:
                with self.monitor.track_time(label):
                    action()
        self.monitor.advance_step(t0, t1)
