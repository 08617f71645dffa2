"""ORACLE (test infrastructure, never a product path): the reference's operator-split monodomain step,
restated on the CPU with NumPy/SciPy.

Each class follows one reference class line by line (control flow, order of copies, time arguments);
dolfinx Functions become plain arrays and the PETSc KSP becomes SciPy ``splu`` (the reference's default
``preonly`` + ``lu``, src/beat/base_model.py:153-157) or the Jacobi-preconditioned CG below (PETSc KSPCG
semantics: zero initial guess, preconditioned residual norm, rtol relative to the right-hand side).

  SplittingSolver      <- src/beat/monodomain_solver.py:26-116
  MonodomainModel      <- src/beat/monodomain_model.py:17-98 + src/beat/base_model.py:73-297
  ODESolver            <- src/beat/odesolver.py:46-79,135-225 (DolfinODESolver, identity projection
                          src/beat/utils.py:52-54)
  Stimulus             <- src/beat/stimulation.py:14-24,264-272

Pinned against the reference's own known-answer tests in tests/test_oracle_known_answers.py
(tests/test_odesolver.py, test_monodomain.py, test_monodomain_solver.py, test_stimulation.py of the
reference).  The cell-model arithmetic (gotranx boundary) is not covered by any reference TEST; it is pinned on the
published Niederer activation times instead (within one dt, tests/test_oracle_niederer.py; see oracle/gen_models.py).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


@dataclass
class Stimulus:
    """I_s(t) * (load vector): expr evaluated at the PDE's theta-point (base_model.py:216-223)."""

    load: np.ndarray
    amplitude: Callable[[float], float]

    @staticmethod
    def window(load: np.ndarray, start: float, end: float, value: float) -> "Stimulus":
        # ufl.conditional(And(ge(time, start), le(time, end)), value, 0)   stimulation.py:270
        return Stimulus(load, lambda t: value if (t >= start and t <= end) else 0.0)


def pcg_petsc(A: sp.csr_matrix, b: np.ndarray, dinv: np.ndarray | None, rtol=1e-5, atol=1e-50, max_it=10000,
              x0: np.ndarray | None = None, norm: str = "preconditioned"):
    """KSPCG as PETSc runs it: returns (x, iterations, residual norm, reason)."""
    n = b.shape[0]
    dinv = np.ones(n) if dinv is None else dinv

    def nrm(r, z):
        if norm == "preconditioned":
            return float(np.sqrt(z @ z))
        if norm == "unpreconditioned":
            return float(np.sqrt(r @ r))
        return float(np.sqrt(abs(r @ z)))

    x = np.zeros(n) if x0 is None else x0.copy()
    r = b - A @ x if x0 is not None else b.copy()
    z = dinv * r
    bnorm = nrm(b, dinv * b)
    ttol = max(rtol * bnorm, atol)
    rnorm = nrm(r, z)
    if rnorm <= ttol:
        return x, 0, rnorm, (3 if rnorm <= atol else 2)
    rz = float(r @ z)
    p = z.copy()
    its = 0
    while True:
        q = A @ p
        alpha = rz / float(p @ q)
        x += alpha * p
        r -= alpha * q
        z = dinv * r
        its += 1
        rnorm = nrm(r, z)
        rz_new = float(r @ z)
        if rnorm <= ttol:
            return x, its, rnorm, (3 if rnorm <= atol else 2)
        if its >= max_it:
            return x, its, rnorm, -3
        beta = rz_new / rz
        rz = rz_new
        p = z + beta * p


class MonodomainModel:
    """theta-rule diffusion step.  A = C_m*Mass + dt*theta*K ; b = (C_m*Mass - dt*(1-theta)*K) v_ + dt*sum I_s."""

    def __init__(self, mass: sp.csr_matrix, stiff: sp.csr_matrix, stimuli=None, C_m: float = 1.0, theta: float = 0.5,
                 solver: str = "lu", rtol: float = 1e-5, atol: float = 1e-50, max_it: int = 10000,
                 x0_previous: bool = False, norm: str = "preconditioned"):
        self.mass, self.stiff = mass.tocsr(), stiff.tocsr()
        self.n = mass.shape[0]
        self.stimuli = list(stimuli or [])
        self.C_m, self.theta = float(C_m), float(theta)
        self.solver, self.rtol, self.atol, self.max_it = solver, rtol, atol, max_it
        self.x0_previous, self.norm = x0_previous, norm
        self.time = 0.0
        self.v_ = np.zeros(self.n)  # monodomain_model.py:52
        self.state = np.zeros(self.n)  # monodomain_model.py:53
        self._timestep = 1.0  # default_timestep, base_model.py:162
        self._update_matrices()
        self.ksp = {"iterations": 0, "residual_norm": 0.0, "reason": 0}
        self.total_iterations = 0

    def assign_previous(self):  # monodomain_model.py:59-60
        self.v_[:] = self.state

    def _update_matrices(self):  # base_model.py:188-194
        dt = self._timestep
        self.A = (self.C_m * self.mass + dt * self.theta * self.stiff).tocsr()
        self.B = (self.C_m * self.mass - dt * (1.0 - self.theta) * self.stiff).tocsr()
        self._lu = None
        self._dinv = 1.0 / self.A.diagonal()

    def _rhs(self) -> np.ndarray:  # base_model.py:196-206 with L from monodomain_model.py:90-96
        b = self.B @ self.v_
        for s in self.stimuli:
            a = s.amplitude(self.time)
            if a != 0.0:
                b = b + self._timestep * a * s.load
        return b

    def step(self, interval):  # base_model.py:208-245
        t0, t1 = interval
        dt = t1 - t0
        t = t0 + self.theta * dt
        self.time = t
        if not abs(dt - self._timestep) < 1.0e-12:
            self._timestep = dt
            self._update_matrices()
        b = self._rhs()
        if self.solver == "lu":
            if self._lu is None:
                self._lu = spla.splu(self.A.tocsc())
            self.state[:] = self._lu.solve(b)
            self.ksp = {"iterations": 1, "residual_norm": 0.0, "reason": 4}
        else:
            x, its, rn, reason = pcg_petsc(self.A, b, self._dinv if self.solver == "cg-jacobi" else None, self.rtol,
                                           self.atol, self.max_it, self.v_ if self.x0_previous else None, self.norm)
            self.state[:] = x
            self.ksp = {"iterations": its, "residual_norm": rn, "reason": reason}
            self.total_iterations += its

    def solve(self, interval, dt=None):  # base_model.py:250-297
        T0, T = interval
        if dt is None:
            dt = T - T0
        t0, t1 = T0, T0 + dt
        while True:
            self.step((t0, t1))
            if (t1 + dt) > (T + 1e-12):
                break
            self.assign_previous()
            t0 = t1
            t1 = t0 + dt
        return self.state


@dataclass
class ODESolver:
    """DolfinODESolver with identical P1 spaces for v_ode and v_pde."""

    v_pde: np.ndarray  # the PDE's state array (shared, README.md:173)
    init_states: np.ndarray
    parameters: np.ndarray | None
    fun: Callable
    num_states: int
    v_index: int = 0
    v_ode: np.ndarray = field(default=None)  # type: ignore[assignment]

    def __post_init__(self):  # odesolver.py:148-162
        n = self.v_pde.shape[0]
        if self.v_ode is None:
            self.v_ode = np.zeros(n)
        if np.shape(self.init_states) == (self.num_states, n):
            self.values = np.copy(self.init_states).astype(float)
        else:
            self.values = np.zeros((self.num_states, n))
            self.values.T[:] = self.init_states

    def step(self, t0: float, dt: float):  # odesolver.py:67-79
        self.values[:] = self.fun(states=self.values, t=t0, parameters=self.parameters, dt=dt)

    def to_dolfin(self):  # odesolver.py:164-166
        self.v_ode[:] = self.values[self.v_index, :]

    def from_dolfin(self):  # odesolver.py:168-170
        self.values[self.v_index, :] = self.v_ode

    def ode_to_pde(self):  # odesolver.py:101-107 -> utils.py:52-54
        self.v_pde[:] = self.v_ode

    def pde_to_ode(self):  # odesolver.py:109-115 -> utils.py:52-54
        self.v_ode[:] = self.v_pde


@dataclass
class SplittingSolver:
    pde: MonodomainModel
    ode: ODESolver
    theta: float = 1.0

    def __post_init__(self):  # monodomain_solver.py:33-37
        self.ode.to_dolfin()
        self.ode.ode_to_pde()
        self.pde.assign_previous()

    def solve(self, interval, dt):  # monodomain_solver.py:39-51
        T0, T = interval
        if dt is None:
            dt = T - T0
        t0, t1 = T0, T0 + dt
        while t1 < T + 1e-12:
            self.step((t0, t1))
            t0 = t1
            t1 = t0 + dt

    def step(self, interval):  # monodomain_solver.py:53-116
        theta = self.theta
        t0, t1 = interval
        dt = t1 - t0
        t = t0 + theta * dt
        self.ode.step(t0=t0, dt=theta * dt)
        self.ode.to_dolfin()
        self.ode.ode_to_pde()
        self.pde.assign_previous()
        self.pde.step((t0, t1))
        self.ode.pde_to_ode()
        self.ode.from_dolfin()
        if np.isclose(theta, 1.0):
            self.pde.assign_previous()
        else:
            self.ode.step(t, (1.0 - theta) * dt)
            self.ode.to_dolfin()
            self.ode.ode_to_pde()
            self.pde.assign_previous()
