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
    def to_dolfin(self) -> None: ...

    def from_dolfin(self) -> None: ...

    def ode_to_pde(self) -> None: ...

    def pde_to_ode(self) -> None: ...

    def step(self, t0: float, dt: float) -> None: ...


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

    def solve(self, interval, dt):  # monodomain_solver.py:39-51
        T0, T = interval
        if dt is None:
            dt = T - T0
        t0 = T0
        t1 = T0 + dt
        while t1 < T + EPS:
            self.step((t0, t1))
            t0 = t1
            t1 = t0 + dt

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
        reference's ``solve`` loop for constant stimuli/windows, minus the per-step interpreter cost."""
        if not self._fused:
            raise RuntimeError("solve_on_device needs a device DolfinODESolver sharing the PDE's context")
        self._prepare_fused(t0, t0 + dt)
        self.pde._ctx.split_solve(t0, dt, nsteps, self.theta)
        self.pde.time.value = t0 + (nsteps - 1) * dt + self.pde.parameters["theta"] * dt
        self._mark_fused()

    def _step_protocol(self, interval):
        theta = self.theta
        t0, t1 = interval
        dt = t1 - t0
        t = t0 + theta * dt
        with self.monitor.track_time("total_step"):
            with self.monitor.track_time("ode_step"):
                self.ode.step(t0=t0, dt=theta * dt)
            with self.monitor.track_time("ode_to_dolfin"):
                self.ode.to_dolfin()
            with self.monitor.track_time("ode_to_pde"):
                self.ode.ode_to_pde()
            with self.monitor.track_time("pde_assign_previous_before"):
                self.pde.assign_previous()
            with self.monitor.track_time("pde_step"):
                self.pde.step((t0, t1))
            with self.monitor.track_time("pde_to_ode"):
                self.ode.pde_to_ode()
            with self.monitor.track_time("ode_from_dolfin"):
                self.ode.from_dolfin()
            if np.isclose(theta, 1.0):
                with self.monitor.track_time("pde_assign_previous_after"):
                    self.pde.assign_previous()
            else:
                with self.monitor.track_time("corrective_ode_step"):
                    self.ode.step(t, (1.0 - theta) * dt)
                with self.monitor.track_time("corrective_ode_to_dolfin"):
                    self.ode.to_dolfin()
                with self.monitor.track_time("corrective_ode_to_pde"):
                    self.ode.ode_to_pde()
                with self.monitor.track_time("corrective_pde_assign_previous"):
                    self.pde.assign_previous()
        self.monitor.advance_step(t0, t1)
